// cmix_b200/csrc/ppmd.cuh - the PPMD byte model on the device.
//
// PPMD::ByteUpdate (reference src/models/ppmd.cpp:1328-1338): after every coded byte the model is
// updated with it and emits the distribution of the next byte, which feeds (a) the byte mixer in
// front of the LSTM (byte-mixer.cpp:15-38) and (b) layer-0 input 2076 through a ByteModel read-out
// (byte-model.cpp:8-45). The model itself (ppmd_model.h) is a sequential pointer machine: ONE lane
// per stream walks the suffix chain; the other 31 lanes of its warp only help with the 256-entry
// floor / vocabulary mask / normalisation. It is a producer like the small models: in the compress
// direction it depends on the bytes only, so it runs ahead of everything else on its own CUDA
// stream, one warp per stream, all warps of a launch group in one CTA.
#pragma once
#include "exact_math.h"
#include "ppmd_model.h"
#include "state.h"

namespace cmixb200 {

enum { PPMD_WARPS = 8 };

// probs_[i] = sqp[i]; floor 1; ByteModel::ByteUpdate zeroes symbols outside the vocabulary
// (byte-model.cpp:40-45); probs_ /= probs_.sum() with valarray::sum() adding front to back.
__device__ __forceinline__ void ppmd_emit(const PpmdModel& m, const u8* __restrict__ vocab, float* __restrict__ out,
                                          float* scratch /* shared, 256 */, int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = lane + 32 * k;
    float v = __uint2float_rn(m.sqp[i]);
    if (v < 1.0f) v = 1.0f;
    if (!vocab[i]) v = 0.0f;
    scratch[i] = v;
  }
  __syncwarp();
  float sum = 0.0f;
  if (lane == 0) {
    sum = scratch[0];
#pragma unroll 8
    for (int i = 1; i < 256; ++i) sum = XM_FADD(sum, scratch[i]);
  }
  sum = __shfl_sync(0xffffffffu, sum, 0);
#pragma unroll
  for (int k = 0; k < 8; ++k) { const int i = lane + 32 * k; out[i] = XM_FDIV(scratch[i], sum); }
  __syncwarp();
}

__global__ void ppmd_init_kernel(PpmdModel* m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) ppmd_init(*m);
}

// The model's registers and small tables (15 KB: symbol masks, binary-context and SEE adaptation tables, the
// symbol queue) are hit several times per visited state; in global memory every store makes the next load
// of that line an L2 round trip. Each warp therefore works on a shared-memory copy and writes it back at
// the end of the launch; only the three arenas (contexts, states, text) stay in HBM/L2.
__device__ __forceinline__ void ppmd_copy_words(void* dst, const void* src, int lane) {
  static_assert(sizeof(PpmdModel) % 4 == 0, "PpmdModel is copied in 32-bit words");
  const unsigned* s = reinterpret_cast<const unsigned*>(src);
  unsigned* d = reinterpret_cast<unsigned*>(dst);
  for (int i = lane; i < (int)(sizeof(PpmdModel) / 4); i += 32) d[i] = s[i];
  __syncwarp();
}

struct PpmdWarpShared { PpmdModel m; float scratch[256]; };

// Bulk: warp w of block b serves stream b * PPMD_WARPS + w of the launch group.
__global__ void __launch_bounds__(PPMD_WARPS * 32, 1) ppmd_kernel(const ChunkArgs* __restrict__ args_all, int n_streams) {
  extern __shared__ __align__(16) unsigned char ppmd_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PpmdWarpShared& sh = reinterpret_cast<PpmdWarpShared*>(ppmd_raw)[warp];
  const int s = blockIdx.x * PPMD_WARPS + warp;
  if (s >= n_streams) return;
  const ChunkArgs a = args_all[s];
  if (a.ppmd_gen == nullptr) return;
  StreamState* st = a.st;
  ppmd_copy_words(&sh.m, st->ppmd, lane);
  const u8* vocab = st->small.vocab;
  for (u32 pos = 0; pos < a.n_bytes; ++pos) {
    if (lane == 0) {
      ppmd_update_byte(sh.m, a.bytes[pos]);
      ppmd_prepare_byte(sh.m);
    }
    __syncwarp();
    ppmd_emit(sh.m, vocab, a.ppmd_gen + (size_t)pos * 256, sh.scratch, lane);
  }
  ppmd_copy_words(st->ppmd, &sh.m, lane);
}

// Lock-step: one byte.
__global__ void __launch_bounds__(32, 1) ppmd_byte_kernel(StreamState* st, u32 byte, float* out) {
  extern __shared__ __align__(16) unsigned char ppmd_raw[];
  PpmdWarpShared& sh = *reinterpret_cast<PpmdWarpShared*>(ppmd_raw);
  const int lane = threadIdx.x;
  ppmd_copy_words(&sh.m, st->ppmd, lane);
  if (lane == 0) {
    ppmd_update_byte(sh.m, (int)byte);
    ppmd_prepare_byte(sh.m);
  }
  __syncwarp();
  ppmd_emit(sh.m, st->small.vocab, out, sh.scratch, lane);
  ppmd_copy_words(st->ppmd, &sh.m, lane);
}

}  // namespace cmixb200
