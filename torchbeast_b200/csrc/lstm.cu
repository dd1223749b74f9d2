#include "lstm.cuh"
namespace tb {
size_t lstm_ws_bytes(int64_t, int64_t, int, int, int) { return 256; }
LstmWs lstm_ws(void*, int64_t, int64_t, int, int, int) { return LstmWs(); }
int lstm_forward(const float*, const float*, const float*, const float*, const LstmParams&, int64_t, int64_t, int, int,
                 int, LstmWs&, float*, float*, float*, float*, cudaStream_t) {
  set_error("lstm_forward: not built yet");
  return 3;
}
int lstm_backward(const float*, const float*, const float*, const LstmParams&, const LstmGrads&, int64_t, int64_t, int,
                  int, int, LstmWs&, float*, float*, float*, cudaStream_t) {
  set_error("lstm_backward: not built yet");
  return 3;
}
}  // namespace tb
