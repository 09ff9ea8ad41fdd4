// synth.cu -- deterministic synthetic inputs generated on the device:
// blocks-v1 / uniform-v1 frames and splitmix64 descriptors.  Integer-only, so the device output
// is bit-identical to the oracle's generators; used by bench.py and the full-size tests so that
// large workloads never cross PCIe.
#include "orbx_kernels.cuh"
#include "orbx_math.cuh"

namespace orbx {

constexpr int kSynTW = 64, kSynTH = 16;  // CTA tile
constexpr int kSynMaxRects = 512;        // rectangles that can touch one tile (R = w*h/900 total)

// blocks-v1: ramp background, R rectangles painted in order (the last one covering a pixel
// wins), then +-3 noise.  Each CTA first filters the rectangles that touch its tile.
__global__ void __launch_bounds__(256) k_synth_blocks(uint8_t* __restrict__ dst, int w, int h, size_t row_stride,
                                                      size_t frame_stride, uint64_t seed, uint64_t first_frame,
                                                      int shift_x, uint64_t noise_seed) {
  __shared__ int rx0[kSynMaxRects], ry0[kSynMaxRects], rx1[kSynMaxRects], ry1[kSynMaxRects];
  __shared__ uint8_t rv[kSynMaxRects];
  __shared__ int order[kSynMaxRects];
  __shared__ int n_hit;
  const uint64_t frame = first_frame + blockIdx.z;
  const uint64_t base = splitmix64((seed << 32) ^ frame);
  const uint64_t nbase = splitmix64((noise_seed << 32) ^ frame);
  const int tx0 = blockIdx.x * kSynTW, ty0 = blockIdx.y * kSynTH;
  const int R = (w * h) / 900;
  if (threadIdx.x == 0) n_hit = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < R; k += blockDim.x) {
    const int x0 = (int)(splitmix64(base ^ (uint64_t)(5 * k + 1)) % (uint64_t)w) - shift_x;
    const int y0 = (int)(splitmix64(base ^ (uint64_t)(5 * k + 2)) % (uint64_t)h);
    const int rw = 8 + (int)(splitmix64(base ^ (uint64_t)(5 * k + 3)) % 82);
    const int rh = 8 + (int)(splitmix64(base ^ (uint64_t)(5 * k + 4)) % 82);
    if (x0 < tx0 + kSynTW && x0 + rw > tx0 && y0 < ty0 + kSynTH && y0 + rh > ty0) {
      const int j = atomicAdd(&n_hit, 1);
      if (j < kSynMaxRects) {
        rx0[j] = x0; ry0[j] = y0; rx1[j] = x0 + rw; ry1[j] = y0 + rh;
        rv[j] = (uint8_t)(splitmix64(base ^ (uint64_t)(5 * k + 5)) % 256);
        order[j] = k;
      }
    }
  }
  __syncthreads();
  const int nh = min(n_hit, kSynMaxRects);
  uint8_t* frame_ptr = dst + (size_t)blockIdx.z * frame_stride;
  for (int i = threadIdx.x; i < kSynTW * kSynTH; i += blockDim.x) {
    const int x = tx0 + (i % kSynTW), y = ty0 + (i / kSynTW);
    if (x >= w || y >= h) continue;
    int v = 40 + (160 * x) / (w - 1);
    int best = -1;
    for (int j = 0; j < nh; j++)
      if (order[j] > best && x >= rx0[j] && x < rx1[j] && y >= ry0[j] && y < ry1[j]) { best = order[j]; v = rv[j]; }
    const int nz = (int)(splitmix64(nbase ^ 0xABCDEFull ^ ((uint64_t)y << 20) ^ (uint64_t)x) % 7) - 3;
    v += nz;
    frame_ptr[(size_t)y * row_stride + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

__global__ void __launch_bounds__(256) k_synth_uniform(uint8_t* __restrict__ dst, int w, int h, size_t row_stride,
                                                       size_t frame_stride, uint64_t seed, uint64_t first_frame) {
  const uint64_t frame = first_frame + blockIdx.z;
  const uint64_t base = splitmix64((seed << 32) ^ frame);
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  dst[(size_t)blockIdx.z * frame_stride + (size_t)y * row_stride + x] =
      (uint8_t)(splitmix64(base ^ ((uint64_t)y << 20) ^ (uint64_t)x) & 255);
}

int launch_synth(int kind, uint8_t* dst, int frames, int w, int h, size_t row_stride, size_t frame_stride,
                 uint64_t seed, uint64_t first_frame, int shift_x, uint64_t noise_seed, cudaStream_t st) {
  // gridDim.z is limited to 65535 frames per launch
  int launches = 0;
  for (int f0 = 0; f0 < frames; f0 += 32768) {
    const int nf = frames - f0 < 32768 ? frames - f0 : 32768;
    uint8_t* d = dst + (size_t)f0 * frame_stride;
    if (kind == 0) {
      dim3 grid((w + kSynTW - 1) / kSynTW, (h + kSynTH - 1) / kSynTH, nf);
      k_synth_blocks<<<grid, 256, 0, st>>>(d, w, h, row_stride, frame_stride, seed, first_frame + f0, shift_x, noise_seed);
    } else {
      dim3 grid((w + 255) / 256, h, nf);
      k_synth_uniform<<<grid, 256, 0, st>>>(d, w, h, row_stride, frame_stride, seed, first_frame + f0);
    }
    launches++;
  }
  return launches;
}

// 64-bit word j of row i is splitmix64(seed ^ (4*(first+i)+j)), little endian
__global__ void __launch_bounds__(256) k_synth_desc(uint64_t* __restrict__ dst, int64_t first, int64_t n_words, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_words) dst[i] = splitmix64(seed ^ (uint64_t)(4 * first + i));
}

int launch_synth_descriptors(uint8_t* dst, int64_t first, int64_t n, uint64_t seed, cudaStream_t st) {
  if (n <= 0) return 0;
  const int64_t words = 4 * n;
  k_synth_desc<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint64_t*>(dst), first, words, seed);
  return 1;
}

}  // namespace orbx
