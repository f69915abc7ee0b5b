// cmix_b200/csrc/fxcm.cuh — the resident FXCM model on the device.
//
// One CTA of 4 warps per stream runs fxcm_model.h's phases for every bit of a sub-chunk. FXCM is a PRODUCER like the
// small models: in the compress direction it depends on the coded bytes and on the LSTM's bit read-out only
// (lstmpr / lstmex, reference predictor.cpp:462-466), never on the final mixer, so it runs ahead of the mixer on its
// own CUDA stream and hands over 431 12-bit codes per bit through the `ext` scratch the mixer stages from.
//
// Lane roles in the unit phase (one lane per unit, unit kinds on different warps so that they do not serialise
// each other): warp 0 lanes 0-15 = context maps 0-15, warp 1 lanes 0-14 = maps 16-30, warp 2 lanes 0-6 = the
// stationary maps, lane 7 = run map, warp 3 lane 0 = match model 2, lane 1 = sparse match model. The int16 mixers
// (SGD and dot products) use all 128 lanes; their integer sums are exact under any association.
// The mutable scalar state (9 KB) and the text-analysis state (40 KB) live in shared memory for the launch.
#pragma once
#include "exact_math.h"
#include "fxcm_model.h"
#include "bytemodel.cuh"
#include "state.h"

namespace cmixb200 {

enum { FX_THREADS = 128 };

#ifdef FX_PROF
__device__ unsigned long long g_fx_prof[2][24];
#define FX_T(k) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_fx_prof[sh.prof_row][k], (unsigned long long)(now_ - sh.prof_t)); sh.prof_t = now_; } } while (0)
#else
#define FX_T(k) do { } while (0)
#endif

struct FxShared {
  fx::State S;
  fx::TextState X;
  int part[4][10];
  int dots[10];
#ifdef FX_PROF
  long long prof_t; int prof_row;
#endif
};

__device__ __forceinline__ int fx_unit_of(int tid) {
  if (tid < 16) return fx::U_MAP0 + tid;
  if (tid >= 32 && tid < 47) return fx::U_MAP0 + 16 + (tid - 32);
  if (tid >= 64 && tid < 71) return tid - 64;
  if (tid == 71) return fx::U_RCM;
  if (tid == 96) return fx::U_MATCH;
  if (tid == 97) return fx::U_SMATCH;
  return -1;
}

__device__ __forceinline__ void fx_copy_words(void* dst, const void* src, size_t bytes, int tid) {
  u32* d = (u32*)dst; const u32* s = (const u32*)src;
  for (size_t i = tid; i < bytes / 4; i += FX_THREADS) d[i] = s[i];
}
static_assert(sizeof(fx::State) % 4 == 0 && sizeof(fx::TextState) % 4 == 0, "state blocks are copied word by word");

__device__ __forceinline__ void fx_load(FxShared& sh, fx::State* g, int tid) {
  fx::TextState* gx = g->text;
  fx_copy_words(&sh.S, g, sizeof(fx::State), tid);
  fx_copy_words(&sh.X, gx, sizeof(fx::TextState), tid);
  __syncthreads();
  if (tid == 0) sh.S.text = &sh.X;
  __syncthreads();
}
__device__ __forceinline__ void fx_store(FxShared& sh, fx::State* g, fx::TextState* gx, int tid) {
  __syncthreads();
  if (tid == 0) sh.S.text = gx;
  __syncthreads();
  fx_copy_words(g, &sh.S, sizeof(fx::State), tid);
  fx_copy_words(gx, &sh.X, sizeof(fx::TextState), tid);
}

// One bit: FXCM::Perceive(bit) (fxcmv1.cpp:4909-4912 -> update1 :4758). All FX_THREADS lanes call it.
__device__ void fx_bit(FxShared& sh, int y, int lstmpr, int lstmex, int tid) {
  fx::State& S = sh.S;
#ifdef FX_PROF
  if (tid == 0) { sh.prof_t = clock64(); sh.prof_row = (S.bpos == 7) ? 0 : 1; }
#endif
  if (tid == 0) fx::bit_head(S, y, lstmpr, lstmex);
  __syncthreads();
  FX_T(0);
  fx::bit_train(S, tid, FX_THREADS);
  __syncthreads();
  FX_T(1);
  if (tid == 0) fx::bit_prepare(S);
  __syncthreads();
  FX_T(2);
  const int u = fx_unit_of(tid);
  if (u >= 0) fx::bit_unit(S, u);
  __syncthreads();
  FX_T(3);
  if (tid == 0) fx::bit_select(S);
  __syncthreads();
  FX_T(4);
  {
    int part[10];
    fx::bit_dot_partial(S, tid, FX_THREADS, part);
#pragma unroll
    for (int i = 0; i < 10; ++i) part[i] = __reduce_add_sync(0xffffffffu, part[i]);
    if ((tid & 31) == 0) for (int i = 0; i < 10; ++i) sh.part[tid >> 5][i] = part[i];
  }
  __syncthreads();
  FX_T(5);
  if (tid == 0) {
    int dots[10];
    for (int i = 0; i < 10; ++i) dots[i] = sh.part[0][i] + sh.part[1][i] + sh.part[2][i] + sh.part[3][i];
    fx::bit_tail(S, dots);
  }
  __syncthreads();
  FX_T(6);
}

// Bulk: CTA b serves stream b of the launch group. Writes ext[t][0..430] for every bit t of the sub-chunk (the codes
// the model holds BEFORE perceiving bit t) and, when a.ext_replay is given, copies the non-resident PAQ8 slots next to them.
__global__ void __launch_bounds__(FX_THREADS, 1) fxcm_kernel(const ChunkArgs* __restrict__ args_all) {
  extern __shared__ __align__(16) unsigned char fx_raw[];
  FxShared& sh = *reinterpret_cast<FxShared*>(fx_raw);
  const ChunkArgs a = args_all[blockIdx.x];
  if (a.fx == nullptr) return;
  const int tid = threadIdx.x;
  fx::TextState* gx = a.fx->text;
  fx_load(sh, a.fx, tid);
  const u32 n_bits = a.n_bytes * 8;
  for (u32 t = 0; t < n_bits; ++t) {
    if (!a.pretrain) {
      u16* out = a.ext_gen + (size_t)t * N_EXT;
      for (int k = tid; k < fx::N_OUT; k += FX_THREADS) out[k] = sh.S.codes[k];
      if (a.ext_replay && !a.paq8) for (int k = fx::N_OUT + tid; k < N_EXT; k += FX_THREADS) out[k] = a.ext_replay[(size_t)t * N_EXT + k];
    }
    const int y = (a.bytes[t >> 3] >> (7 - (t & 7))) & 1;
    int lstmpr = 0, lstmex = 0;
    if (!a.pretrain) { const u32 v = a.lstm_fx[t]; lstmpr = (int)(v & 0xffff); lstmex = (int)(v >> 16); }
    else { lstmpr = sh.S.lstmpr; lstmex = sh.S.lstmex; }
    fx_bit(sh, y, lstmpr, lstmex, tid);
  }
  fx_store(sh, a.fx, gx, tid);
}

// Lock-step: one bit per launch (the decoder's order), queued behind the mixer / LSTM update of the same bit. Lane 0 first
// takes the LSTM's read-out of the NEXT bit (ByteMixer::Predict + Discretize, predictor.cpp:462-465); the codes for the next
// Predict() land in ext_bit[0..430].
__global__ void __launch_bounds__(FX_THREADS, 1) fxcm_bit_kernel(StreamState* st, fx::State* g, int y, int pretrain, u16* ext_bit) {
  extern __shared__ __align__(16) unsigned char fx_raw[];
  FxShared& sh = *reinterpret_cast<FxShared*>(fx_raw);
  const int tid = threadIdx.x;
  fx::TextState* gx = g->text;
  fx_load(sh, g, tid);
  if (tid == 0) {
    if (pretrain) { sh.dots[0] = sh.S.lstmpr; sh.dots[1] = sh.S.lstmex; }
    else {
      const ByteModelState& b = st->lstm.bm;
      int ex;
      const float p = bytemodel_predict(b.probs, b.bot, b.top, &ex);
      sh.dots[0] = (int)(u32)XM_FADD(1.0f, XM_FMUL(4094.0f, p));
      sh.dots[1] = ex;
    }
  }
  __syncthreads();
  const int lstmpr = sh.dots[0], lstmex = sh.dots[1];
  __syncthreads();
  fx_bit(sh, y, lstmpr, lstmex, tid);
  if (ext_bit) for (int k = tid; k < fx::N_OUT; k += FX_THREADS) ext_bit[k] = sh.S.codes[k];
  fx_store(sh, g, gx, tid);
}

}  // namespace cmixb200
