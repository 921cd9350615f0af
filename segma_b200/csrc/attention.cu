// C entry points of the fused self-attention (the kernel lives in attention_tc5.cu: tcgen05 / TMEM, sm_100a).
// There is one backend: inputs the tensor-memory kernel cannot take (a position-bias table whose rows are not 16-byte
// aligned) are rejected, not routed to another implementation.
#include "common.cuh"

namespace segma {

int launch_attention_tc5(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                         const float* pos_bias, int pb_ld, int bias_mode, void* out, cudaStream_t st);

}  // namespace segma

using namespace segma;

extern "C" {

int segma_attention(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                    const float* pos_bias, int pos_bias_ld, void* out, void* stream) {
  SEGMA_REQUIRE(n_windows >= 0 && T > 0 && n_heads > 0 && n_query >= 0 && n_query <= T, "segma_attention: bad shape");
  if (n_windows == 0 || n_query == 0) return SEGMA_OK;
  SEGMA_REQUIRE(qkv && out, "segma_attention: NULL buffer");
  SEGMA_REQUIRE((gate == nullptr) == (pos_bias == nullptr), "segma_attention: gate and pos_bias go together");
  SEGMA_REQUIRE(n_heads <= 65535 && n_windows <= 65535, "segma_attention: grid too large");
  SEGMA_REQUIRE(pos_bias == nullptr || pos_bias_ld >= T, "segma_attention: pos_bias_ld %d < T %d", pos_bias_ld, T);
  SEGMA_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "segma_attention: qkv must be 16-byte aligned");
  SEGMA_REQUIRE(pos_bias == nullptr || (pos_bias_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(pos_bias) & 15) == 0),
                "segma_attention: pos_bias rows must be 16-byte aligned (pos_bias_ld a multiple of 4 floats)");
  return launch_attention_tc5(qkv, n_windows, T, n_heads, n_query, gate, pos_bias, pos_bias_ld, pos_bias ? 1 : 0, out,
                              (cudaStream_t)stream);
}

int segma_attention_rel(const void* qkv, int n_windows, int T, int n_heads, int n_query, const float* gate,
                        const float* rel_bias, void* out, void* stream) {
  SEGMA_REQUIRE(n_windows >= 0 && T > 0 && n_heads > 0 && n_query >= 0 && n_query <= T, "segma_attention_rel: bad shape");
  if (n_windows == 0 || n_query == 0) return SEGMA_OK;
  SEGMA_REQUIRE(qkv && out && gate && rel_bias, "segma_attention_rel: NULL buffer");
  SEGMA_REQUIRE(n_heads <= 65535 && n_windows <= 65535, "segma_attention_rel: grid too large");
  SEGMA_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "segma_attention_rel: qkv must be 16-byte aligned");
  return launch_attention_tc5(qkv, n_windows, T, n_heads, n_query, gate, rel_bias, 0, 2, out, (cudaStream_t)stream);
}

}  // extern "C"
