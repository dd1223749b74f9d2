"""Empty stand-in for the reference's C++ extension `libtorchbeast`.

polybeast_learner imports it at module import but only train() touches it.
TEST INFRASTRUCTURE ONLY.
"""
