// Weight stream and launcher of the tensor-core (tcgen05 / TMEM) encoder, encode_tc.cu.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "encode_tc_stream.hpp"
#include "model.cuh"

namespace vqvdb {

cudaError_t configure_encode_tc();
// tap_stage >= 0 additionally writes an fp32 activation to tap_out (bring-up aid; -1 in production):
//   0 pre (GN+ReLU) [n][16][512], 6 res16 conv1 out [n][16][512], 1 res16 block out [n][16][512], 2 down out [n][32][64],
//   7 res32 conv1 out [n][32][64], 3 res32 block out, 4 after attention [n][32][64], 5 z [n][128][64];
//   100 = cycle counters (see tools/enc_tc_prof.py).
cudaError_t launch_encode_tc(const EncoderWeights& w, const EncoderTcStream& stream_tab, const float* dev_leaves, int64_t n_leaves,
                             uint8_t* dev_indices, int num_sms, cudaStream_t stream, int tap_stage = -1, float* tap_out = nullptr);

}  // namespace vqvdb
