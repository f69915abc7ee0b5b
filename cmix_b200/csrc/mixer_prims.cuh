// cmix_b200/csrc/mixer_prims.cuh
//
// Building blocks shared by the warp-specialised gated-mixer kernel (mixer_v3.cuh): named barriers,
// NCCL-"LL"-style {value, sequence} slots for cross-CTA hand-off without cluster fences, the TMA
// (cp.async.bulk + mbarrier) helpers that move whole 8.4 KB weight rows, input staging, and the
// per-phase cycle accounting used by tools/gpu_prof.py. mixer.cuh documents the parity rules and holds
// the lock-step / barrier-per-phase kernels.
#pragma once
#include "mixer.cuh"

namespace cmixb200 {

// Warp roles. The SM's 4 schedulers own warps (w % 4); the arbiter prefers the highest warp id,
// so the latency-critical chain warp (15) and the tail warp (14) each get a scheduler of their own:
// movers are the warps with (w % 4) < 2; warps 2, 3, 6, 7, 10, 11 stay parked at the final barrier.
enum { V2_NBUF = 20, V2_RING = 4, V2_M_WARPS = 8, V2_M_THREADS = V2_M_WARPS * 32, V2_CM_THREADS = V2_M_THREADS + 32,
       V2_C_WARP = 15, V2_T_WARP = 14 };

struct RowJob { int buf; int mixer; u32 load_slot; u32 evict_slot; int do_evict; int do_load; };

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void spin_until_ge(volatile u32* p, u32 v) {
  while (*p < v) { }
}

// "LL" message slots (as in NCCL's low-latency protocol): a float and its sequence number are written
// with one 8-byte store, so the consumer needs no fence - it polls the slot until the sequence matches.
// (A cluster-scope fence costs an L1 invalidate + a drain of the warp's outstanding global stores.)
__device__ __forceinline__ void ll_store(uint2* slot, float v, u32 seq) {
  *reinterpret_cast<volatile unsigned long long*>(slot) = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
}
__device__ __forceinline__ float ll_wait(const uint2* slot, u32 seq) {
  unsigned long long m;
  do { m = *reinterpret_cast<const volatile unsigned long long*>(slot); } while ((u32)(m >> 32) != seq);
  return __uint_as_float((u32)m);
}

enum { BAR_READY = 1, BAR_COEFF = 2, BAR_MOVERS = 3 };
// Per-phase cycle accounting (debug). BAR.SYNC does not block at issue, so a clock read placed right
// after a barrier would capture the issue time; a dependent shared-memory load + MOV in front of the
// clock read makes the sample wait for the barrier's release.
#define V2_PROF(cond, slot) do { if (cond) { \
    unsigned dummy_ = *reinterpret_cast<volatile unsigned*>(&sh.n_jobs), sink_; \
    asm volatile("mov.u32 %0, %1;" : "=r"(sink_) : "r"(dummy_)); \
    const long long now_ = clock64(); pacc[(slot) & 7] += (unsigned long long)(now_ - tprev) + (sink_ & 0u); tprev = now_; } } while (0)
#define V2_PROF_DUMP(cond, base, n) do { if (cond) { for (int q_ = 0; q_ < (n); ++q_) a.prof[(base) + q_] += pacc[((base) + q_) & 7]; } } while (0)

__device__ __forceinline__ bool selector_is_bit_level(int sel) {
  return sel == S_AUX || sel == S_LONGBIT || (sel >= S_BC0 && sel <= S_BC_RB1);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- TMA (bulk async copy) helpers: one instruction moves a whole 8.4 KB weight row ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W;\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_row(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_row(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

// Stage the 2078 layer-0 inputs with all global loads issued before any use (2 round trips).
template <int NT>
__device__ __forceinline__ void stage_inputs_v2(float* x, const float* lut, const u16* ext, const float* small_x,
                                                float lstm_x, int mtid) {
  enum { PER = (N_INPUTS + NT - 1) / NT };
  u32 code[PER]; float direct[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int k = mtid + q * NT;
    code[q] = 0x10000u; direct[q] = 0.0f;
    if (k < N_INPUTS) {
      if (k < 3) direct[q] = small_x[k];
      else if (k < 3 + N_EXT) code[q] = ext ? (u32)__ldcs(&ext[k - 3]) : 0xFFFFu;
      else if (k < 2076) direct[q] = small_x[k - N_EXT];
      else if (k == 2076) direct[q] = small_x[N_SMALL];
      else direct[q] = lstm_x;
    }
  }
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int k = mtid + q * NT;
    if (k < N_INPUTS) x[k] = code[q] == 0x10000u ? direct[q] : lut[code[q] == 0xFFFFu ? 4096 : (code[q] > 4095u ? 4095u : code[q])];
  }
}

}  // namespace cmixb200
