"""CPU: host side of the learner-queue ingest (N1): slot carving and the native per-actor column writer
(tb_host_write_rollout_column - pure host memcpy, no GPU needed)."""
import ctypes

import torch


def test_slot_layout_is_aligned_and_disjoint():
    from torchbeast_b200 import staging
    spec = staging.spec_for(80, 32, 6)
    offs, total = staging._carve(spec)
    end = 0
    for name, (off, nbytes) in offs.items():
        assert off % 256 == 0 and off >= end, name
        end = off + nbytes
    assert total >= end and offs["frame"][1] == 81 * 32 * 4 * 84 * 84
    # frames dominate: 73.2 MB of the ~73.4 MB a full slot moves per step
    assert 73_150_000 < sum(n for _, n in offs.values()) < 73_500_000


def test_native_column_writer_places_every_leaf():
    from torchbeast_b200 import _lib, staging
    T, B, A = 3, 4, 6
    spec = staging.spec_for(T, B, A, use_last_action=False)
    offs, total = staging._carve(spec)
    raw = torch.zeros(total, dtype=torch.uint8)
    names = list(spec)
    T1 = T + 1
    rolls = []
    for b in (2, 0):
        roll = {k: (torch.rand(*((T1,) + spec[k][0][2:])) * 100).to(spec[k][1]) for k in names}
        n = len(names)
        o = (ctypes.c_int64 * n)(*[offs[k][0] for k in names])
        r = (ctypes.c_int64 * n)(*[offs[k][1] // (T1 * B) for k in names])
        srcs = (ctypes.c_void_p * n)(*[roll[k].data_ptr() for k in names])
        rc = _lib.lib().tb_host_write_rollout_column(ctypes.c_void_p(raw.data_ptr()), ctypes.cast(o, ctypes.c_void_p),
                                                     ctypes.cast(r, ctypes.c_void_p), n, T1, B, b, ctypes.cast(srcs, ctypes.c_void_p))
        assert rc == 0
        rolls.append((b, roll))
    for k in names:
        off, nb = offs[k]
        v = raw[off:off + nb].view(spec[k][1]).view(spec[k][0])
        for b, roll in rolls:
            assert torch.equal(v[:, b], roll[k]), k
        assert float(v[:, 1].float().abs().sum()) == 0 and float(v[:, 3].float().abs().sum()) == 0
    h = _lib.lib()
    assert h.tb_host_write_rollout_column(None, None, None, 1, 1, 1, 0, None) != 0 and b"null pointer" in h.tb_last_error()
