// placeholder until the packed encode/decode kernels land (next commit)
#include "common.cuh"
extern "C" {
int smaq_packed_layout_for(int64_t, int32_t, int32_t, smaq_packed_layout*) {
  return smaq::fail(SMAQ_ERR_UNSUPPORTED, "packed stream: not built yet");
}
int smaq_encode(const float*, int64_t, const float*, const float*, const smaq_codec_params*, void*, size_t, void*,
                size_t, smaq_stream_t) {
  return smaq::fail(SMAQ_ERR_UNSUPPORTED, "packed stream: not built yet");
}
int smaq_decode(const void*, size_t, int64_t, int32_t, int32_t, int32_t, float*, smaq_stream_t) {
  return smaq::fail(SMAQ_ERR_UNSUPPORTED, "packed stream: not built yet");
}
}
