// k_glue.cu -- the two pieces of the sampler's inner loop around env.step that are neither MLP nor physics, batched:
// the observation normaliser (ZFilter / RunningStat, uhc/khrylib/utils/zfilter.py:7-73) and the sample logger
// (LoggerRL.step / end_episode, uhc/khrylib/rl/core/logger_rl.py:28-44).  HBM-bound streaming kernels: one pass over the
// [N, dim] observations for the batch moments, one pass to normalise.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>
#include "../../include/hoic.h"

namespace {
thread_local std::string g_glue_err;
int glue_err(const char* what, cudaError_t e) { g_glue_err = std::string(what) + ": " + cudaGetErrorString(e); return -1; }

#define ZF_COLS 32       // columns per CTA (one warp-wide coalesced row segment)
#define ZF_ROWS 8        // row lanes per CTA
#define ZF_CHUNKS 32     // row chunks (grid.y) of the moments pass: 20 x 32 CTAs for 617 columns (8 chunks: 20.5 us at 4096 rows)
// Batch moments about the batch's first row c (a sample of the distribution is a good shift: sums of (x - c) and (x - c)^2
// add across CTAs without cancellation, whatever the running mean is), double accumulation.  The last CTA to finish merges the batch into the running statistics (Chan et al.,
// exactly what a sequence of RunningStat.push calls converges to) and into the since-last-sync accumulators, publishes
// float mean / 1 / (std + 1e-8) for the normalising pass and clears the scratch for the next call.
__global__ void __launch_bounds__(ZF_COLS * ZF_ROWS) k_zf_moments(const float* __restrict__ x, int n, int dim, int row_chunks, double* st_n,
                                                                  double* mean, double* m2, double* dn, double* dmean, double* dm2,
                                                                  double* scratch /*[ZF_CHUNKS * 2 dim + 1]*/, float* shift_scale /*[2 dim]*/, int update) {
  const int col = blockIdx.x * ZF_COLS + (threadIdx.x & (ZF_COLS - 1)), rl = threadIdx.x / ZF_COLS;
  __shared__ double s1[ZF_ROWS][ZF_COLS], s2[ZF_ROWS][ZF_COLS];
  __shared__ int last;
  double a1 = 0, a2 = 0;
  if (update && col < dim) {
    const double c = (double)x[col];        // shift = the batch's first row (a sample of the distribution)
    const int rows_per = (n + row_chunks - 1) / row_chunks, r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    for (int r = r0 + rl; r < r1; r += ZF_ROWS) {
      const double d = (double)x[(size_t)r * dim + col] - c;
      a1 += d; a2 += d * d;
    }
  }
  s1[rl][threadIdx.x & (ZF_COLS - 1)] = a1; s2[rl][threadIdx.x & (ZF_COLS - 1)] = a2;
  __syncthreads();
  if (update && rl == 0 && col < dim) {
    for (int k = 1; k < ZF_ROWS; k++) { a1 += s1[k][threadIdx.x]; a2 += s2[k][threadIdx.x]; }
    // one slot per row chunk, summed in a fixed order by the last CTA: the statistics are bitwise reproducible
    scratch[(size_t)blockIdx.y * 2 * dim + col] = a1;
    scratch[(size_t)blockIdx.y * 2 * dim + dim + col] = a2;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    last = atomicAdd((unsigned long long*)(scratch + (size_t)ZF_CHUNKS * 2 * dim), 1ull) == (unsigned long long)(total - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const double nb = (double)n;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    if (update) {
      double S1 = 0, S2 = 0;
      for (int k = 0; k < row_chunks; k++) { S1 += __ldcg(scratch + (size_t)k * 2 * dim + c); S2 += __ldcg(scratch + (size_t)k * 2 * dim + dim + c); }
      const double m0 = mean[c];
      const double bmean = (double)x[c] + S1 / nb, bm2 = fmax(S2 - S1 * S1 / nb, 0.0);
      {
        const double n0 = *st_n, tot = n0 + nb, delta = bmean - m0, w = nb / fmax(tot, 1.0);
        mean[c] = m0 + delta * w;
        m2[c] = m2[c] + bm2 + delta * delta * (n0 * w);
      }
      if (dn) {
        const double n0 = *dn, tot = n0 + nb, delta = bmean - dmean[c], w = nb / fmax(tot, 1.0);
        dmean[c] = dmean[c] + delta * w;
        dm2[c] = dm2[c] + bm2 + delta * delta * (n0 * w);
      }
    }
    const double ntot = *st_n + (update ? nb : 0.0);
    const double var = ntot > 1 ? m2[c] / (ntot - 1) : mean[c] * mean[c];     // RunningStat.var (zfilter.py:36-37)
    shift_scale[c] = (float)mean[c];
    shift_scale[dim + c] = (float)(1.0 / (sqrt(fmax(var, 0.0)) + 1e-8));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (update) { *st_n += nb; if (dn) *dn += nb; }
    ((unsigned long long*)(scratch + (size_t)ZF_CHUNKS * 2 * dim))[0] = 0ull;
  }
}
// y = clip((x - mean) / (std + 1e-8), -clip, clip), the subtraction in double like the reference's numpy arithmetic
__global__ void k_zf_apply(const float* __restrict__ x, size_t total, int dim, const double* __restrict__ mean, const float* __restrict__ shift_scale,
                           float clip, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % dim);
    const float v = (float)(((double)x[i] - mean[c]) * (double)shift_scale[dim + c]);
    y[i] = fminf(fmaxf(v, -clip), clip);
  }
}
// the same pass with a second output for the policy MLP's first GEMM: y16[n][ld16] = bfloat16(y) (round to nearest even, what
// torch.autocast's cast does), columns dim..ld16-1 zero (the MLP pads its 617 inputs to a multiple of 64, ppo.MLP)
__global__ void k_zf_apply16(const float* __restrict__ x, int n, int dim, int ld16, const double* __restrict__ mean,
                             const float* __restrict__ shift_scale, float clip, float* __restrict__ y, __nv_bfloat16* __restrict__ y16) {
  const size_t total = (size_t)n * ld16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld16);
    const size_t r = i / ld16;
    float v = 0.f;
    if (c < dim) {
      v = (float)(((double)x[r * dim + c] - mean[c]) * (double)shift_scale[dim + c]);
      v = fminf(fmaxf(v, -clip), clip);
      y[r * dim + c] = v;
    }
    y16[i] = __float2bfloat16_rn(v);
  }
}
// PolicyGaussian.select_action for the batch with the sampler's mean_action switch (agent_handmimic.py:463-466, 484):
// explore[i] = u[i] < noise_rate (1 - mean_action), action = mean + exp(log_std) * noise * explore
__global__ void k_sample_action(const __nv_bfloat16* __restrict__ mean16, const float* __restrict__ mean32, const float* __restrict__ log_std,
                                const float* __restrict__ noise, const float* __restrict__ u, const float* __restrict__ noise_rate, int n,
                                int act_dim, float* __restrict__ action, float* __restrict__ explore) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * act_dim) return;
  const int r = i / act_dim, c = i - r * act_dim;
  const float e = u[r] < *noise_rate ? 1.f : 0.f;
  const float mu = mean16 ? __bfloat162float(mean16[i]) : mean32[i];
  action[i] = mu + expf(log_std[c]) * noise[i] * e;
  if (c == 0) explore[r] = e;
}

// LoggerRL for a batch: sums[0..12] = steps, episodes, c_reward, info[9], finished-episode reward; minmax = c_reward min / max,
// episode reward min / max; episode_reward[N] running sums (reset where an episode ended).  One CTA.
__global__ void __launch_bounds__(1024) k_logger(const float* __restrict__ reward, const float* __restrict__ info, const uint8_t* __restrict__ done, int n,
                                                 double* sums, double* minmax, double* episode_reward) {
  __shared__ double red[32][16];
  double acc[13];
  for (int k = 0; k < 13; k++) acc[k] = 0;
  double rmin = INFINITY, rmax = -INFINITY, emin = INFINITY, emax = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = reward[i];
    const bool d = done[i] != 0;
    acc[0] += 1; acc[1] += d ? 1 : 0; acc[2] += r;
    for (int k = 0; k < 9; k++) acc[3 + k] += (double)info[(size_t)i * 9 + k];
    rmin = fmin(rmin, r); rmax = fmax(rmax, r);
    if (episode_reward) {
      const double ep = episode_reward[i] + r;
      if (d) { acc[12] += ep; emin = fmin(emin, ep); emax = fmax(emax, ep); }
      episode_reward[i] = d ? 0.0 : ep;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 13; k++) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    rmin = fmin(rmin, __shfl_xor_sync(0xffffffffu, rmin, o)); rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o)); emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  }
  __shared__ double mm[32][4];
  if (lane == 0) { mm[warp][0] = rmin; mm[warp][1] = rmax; mm[warp][2] = emin; mm[warp][3] = emax; }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  if (threadIdx.x < 13) {
    double v = 0;
    for (int w = 0; w < nw; w++) v += red[w][threadIdx.x];
    sums[threadIdx.x] += v;
  } else if (threadIdx.x >= 32 && threadIdx.x < 36) {
    const int k = threadIdx.x - 32;
    double v = mm[0][k];
    for (int w = 1; w < nw; w++) v = (k & 1) ? fmax(v, mm[w][k]) : fmin(v, mm[w][k]);
    minmax[k] = (k & 1) ? fmax(minmax[k], v) : fmin(minmax[k], v);
  }
}
}  // namespace

extern "C" const char* hoic_glue_last_error(void) { return g_glue_err.c_str(); }

extern "C" int hoic_zfilter16(const float* obs, int n, int dim, double* st_n, double* mean, double* m2, double* dn, double* dmean, double* dm2,
                              double* scratch, float* shift_scale, float clip, int update, float* x_out, void* x16_out, int ld16, void* stream) {
  if (!obs || n < 1 || dim < 1 || !st_n || !mean || !m2 || !scratch || !shift_scale || !x_out || (x16_out && ld16 < dim)) {
    g_glue_err = "hoic_zfilter: bad arguments";
    return -1;
  }
  cudaStream_t sm = (cudaStream_t)stream;
  const int chunks = update ? ZF_CHUNKS : 1;
  k_zf_moments<<<dim3((dim + ZF_COLS - 1) / ZF_COLS, chunks), ZF_COLS * ZF_ROWS, 0, sm>>>(obs, n, dim, chunks, st_n, mean, m2, dn, dmean, dm2, scratch,
                                                                                       shift_scale, update);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return glue_err("k_zf_moments", e);
  const size_t total = (size_t)n * (x16_out ? ld16 : dim);
  const int grid = (int)((total + 1023) / 1024 < 1184 ? (total + 1023) / 1024 : 1184);
  if (x16_out) k_zf_apply16<<<grid, 256, 0, sm>>>(obs, n, dim, ld16, mean, shift_scale, clip, x_out, (__nv_bfloat16*)x16_out);
  else k_zf_apply<<<grid, 256, 0, sm>>>(obs, total, dim, mean, shift_scale, clip, x_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return glue_err("k_zf_apply", e);
  return 0;
}
extern "C" int hoic_zfilter(const float* obs, int n, int dim, double* st_n, double* mean, double* m2, double* dn, double* dmean, double* dm2,
                            double* scratch, float* shift_scale, float clip, int update, float* x_out, void* stream) {
  return hoic_zfilter16(obs, n, dim, st_n, mean, m2, dn, dmean, dm2, scratch, shift_scale, clip, update, x_out, nullptr, 0, stream);
}
extern "C" int hoic_sample_action(const void* mean, int mean_is_bf16, const float* log_std, const float* noise, const float* u,
                                  const float* noise_rate, int n, int act_dim, float* action, float* explore, void* stream) {
  if (!mean || !log_std || !noise || !u || !noise_rate || n < 1 || act_dim < 1 || !action || !explore) { g_glue_err = "hoic_sample_action: bad arguments"; return -1; }
  k_sample_action<<<(n * act_dim + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mean_is_bf16 ? (const __nv_bfloat16*)mean : nullptr,
                                                                            mean_is_bf16 ? nullptr : (const float*)mean, log_std, noise, u,
                                                                            noise_rate, n, act_dim, action, explore);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return glue_err("k_sample_action", e);
  return 0;
}

extern "C" int hoic_logger_step(const float* reward, const float* reward_info, const uint8_t* done, int n, double* sums, double* minmax,
                                double* episode_reward, void* stream) {
  if (!reward || !reward_info || !done || n < 1 || !sums || !minmax) { g_glue_err = "hoic_logger_step: bad arguments"; return -1; }
  k_logger<<<1, 1024, 0, (cudaStream_t)stream>>>(reward, reward_info, done, n, sums, minmax, episode_reward);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return glue_err("k_logger", e);
  return 0;
}
