// cmix_b200/csrc/producers.h — launchers of the two resident producer models. Their kernels live in translation units of
// their own (fxcm_dev.cu, paq8_dev.cu) so that the three big device programs compile side by side.
#pragma once
#include <cuda_runtime.h>

#include "state.h"

namespace cmixb200 {
namespace fx { struct State; }
namespace p8 { struct State; }

cudaError_t fxcm_configure();                                    // opt-in shared memory sizes; once per device
void fxcm_launch_chunk(const ChunkArgs* d_args, int n_streams, cudaStream_t s);
void fxcm_launch_bit(StreamState* st, fx::State* g, int y, int pretrain, u16* ext_bit, cudaStream_t s, const u32* dbit = nullptr);
cudaError_t paq8_configure();
void paq8_launch_chunk(const ChunkArgs* d_args, int n_streams, cudaStream_t s);
void paq8_launch_bit(p8::State* g, int y, u16* ext_bit, cudaStream_t s, const u32* dbit = nullptr);

}  // namespace cmixb200
