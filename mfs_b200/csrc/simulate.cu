// Data simulators: the step immediately BEFORE the filter.  One trajectory per thread, the
// whole time loop inside the kernel, only ys (and optionally xs / x0) leave the SM.
//
//   1-D  simulate_sde(m_and_cov = tme.mean_and_cov(order), x0, dt, T, key, integration_steps)     mfs/utils.py:190-249
//        as called by benes_bernoulli / well_poisson                                 mfs/one_dim/ss_models.py:49-54, 86-91
//        x0 ~ GaussianSum1D.sampler                                                                   mfs/utils.py:55-58
//        y_k ~ Bernoulli(logistic(x_k)) / Poisson(emission(x_k, p2))      dardel/benes_bernoulli/mf.py:80,
//                                                                         dardel/parameter_estimation/mf.py:65
//   2-D  prey_predator.simulate: Milstein sub-steps of the Lotka--Volterra SDE + Bernoulli(emission(x[0]))
//                                                                                    mfs/multi_dims/ss_models.py:76-93
//        x0 ~ GaussianSumND.sampler                                                                mfs/utils.py:101-105
//
// Random numbers.  The reference consumes jax.random (threefry) keys from rng_keys.npy; that stream cannot be
// reproduced without JAX, so the simulators define their own counter-based stream: Philox4x32-10 (Salmon et al. 2011),
// key = the 64-bit seed, counter = (trajectory id lo, trajectory id hi, time index, draw index).  A trajectory's
// numbers depend only on (seed, global trajectory id), so a batch sharded over ranks with `traj_offset` equals the
// single-GPU batch bit for bit, and the stream can be regenerated on the host from (seed, trajectory id) alone.
//   time index 0: initial condition (draw 0: mixture component, draw 1: normals); time index t+1: step t;
//   draw j < 2^31: sub-step normals (1-D: sub-steps 2j, 2j+1; 2-D: sub-step j, one normal per dimension);
//   draw 0x80000000: the measurement; draw 0x80000001: the sign of the exact Benes transition.
// A Philox block gives two uniforms u = (k + 1/2) 2^-52 (k = 52 random bits) and, by Box--Muller, two normals.
#include <cuda_runtime.h>

#include "common.h"
#include "models.cuh"

namespace mfs {

constexpr uint32_t kDrawMeasurement = 0x80000000u;
constexpr uint32_t kDrawSign = 0x80000001u;

MFS_DEV uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

MFS_DEV double uniform52(uint32_t a, uint32_t b) {
  const uint64_t k = ((uint64_t)a << 20) | (uint64_t)(b >> 12);
  return ((double)k + 0.5) * 2.220446049250313e-16;   // (k + 1/2) 2^-52, exact, in (0, 1)
}

struct Draw { double ua, ub; };

MFS_DEV Draw draw(uint64_t traj, uint32_t time_index, uint32_t j, uint64_t seed) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)traj, (uint32_t)(traj >> 32), time_index, j), (uint32_t)seed,
                                (uint32_t)(seed >> 32));
  Draw d;
  d.ua = uniform52(r.x, r.y);
  d.ub = uniform52(r.z, r.w);
  return d;
}

// Box--Muller: (z0, z1) = sqrt(-2 ln ua) (cos 2 pi ub, sin 2 pi ub)
MFS_DEV void normals(const Draw& d, double& z0, double& z1) {
  const double r = sqrt(-2.0 * log(d.ua));
  double s, c;
  sincospi(2.0 * d.ub, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

MFS_DEV int mixture_component(double u, const double* weights, int K) {
  double acc = 0.0;
  int k = 0;
  for (; k < K - 1; ++k) {
    acc += weights[k];
    if (u < acc) break;
  }
  return k;
}

// y ~ p(. | x) from the measurement draw.  Bernoulli: ua < p.  Poisson: inversion by sequential search on ua for rates
// up to 64, Normal approximation beyond.
// Gaussian: h x + r z0.
MFS_DEV double measure(int meas_id, const double* mp, double x, const Draw& d) {
  if (meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) {
    const double p = 1.0 / (1.0 + exp(-(x * x * x / mp[0] - mp[1])));
    return d.ua < p ? 1.0 : 0.0;
  }
  if (meas_id == MFS_MEAS_POISSON_SOFTPLUS) {
    const double lam = log(1.0 + exp(mp[0] * x));
    if (!(lam <= 64.0)) {
      // large rates: exp(-lam) underflows for lam > 745 and the sequential search would run its full bound with the
      // warp waiting; the Normal approximation with continuity correction, floor(lam + sqrt(lam) z + 1/2), is within
      // O(lam^-1/2) of the Poisson law (jax.random.poisson switches to a rejection sampler there)
      double z0, z1;
      normals(d, z0, z1);
      return fmax(0.0, floor(fma(sqrt(lam), z0, lam) + 0.5));     // NaN rate -> NaN measurement
    }
    double p = exp(-lam), F = p;
    int k = 0;
    while (d.ua > F && k < 1000) {
      ++k;
      p *= lam / (double)k;
      F += p;
    }
    return (double)k;
  }
  double z0, z1;
  normals(d, z0, z1);
  return fma(mp[1], z0, mp[0] * x);
}

// Packs eight uint8 measurements into one 8-byte store when the row is 8-byte aligned and unit-stride in time.
struct YsWriter {
  char* row;
  int64_t stride_t;
  int32_t dtype;
  bool packed;
  uint64_t word;
  MFS_DEV void init(void* ys, int32_t dt, int64_t b, int64_t stride_b, int64_t st) {
    dtype = dt;
    stride_t = st;
    const int64_t esz = dt == MFS_YS_U8 ? 1 : dt == MFS_YS_I32 ? 4 : 8;
    row = reinterpret_cast<char*>(ys) + b * stride_b * esz;
    packed = dt == MFS_YS_U8 && st == 1 && (reinterpret_cast<uintptr_t>(row) & 7) == 0;
    word = 0;
  }
  MFS_DEV void put(int64_t t, int64_t T, double y) {
    if (dtype == MFS_YS_U8) {
      if (packed) {
        word |= (uint64_t)(uint8_t)y << (8 * (t & 7));
        if ((t & 7) == 7) {
          *reinterpret_cast<uint64_t*>(row + (t - 7)) = word;
          word = 0;
        } else if (t + 1 == T) {
          for (int64_t q = t & ~int64_t(7); q <= t; ++q) row[q] = (char)(word >> (8 * (q & 7)));
        }
      } else {
        reinterpret_cast<uint8_t*>(row)[t * stride_t] = (uint8_t)y;
      }
    } else if (dtype == MFS_YS_I32) {
      reinterpret_cast<int32_t*>(row)[t * stride_t] = (int32_t)y;
    } else {
      reinterpret_cast<double*>(row)[t * stride_t] = y;
    }
  }
};

__global__ void __launch_bounds__(128) simulate1d_kernel(const mfs_simulate1d_args P) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const uint64_t traj = (uint64_t)b + P.traj_offset;
  double tp[MFS_MAX_PARAMS], mp[MFS_MAX_PARAMS];
#pragma unroll
  for (int i = 0; i < MFS_MAX_PARAMS; ++i) {
    tp[i] = P.trans_params ? P.trans_params[b * P.trans_param_stride + i] : 0.0;
    mp[i] = P.meas_params ? P.meas_params[b * P.meas_param_stride + i] : 0.0;
  }
  // x0 ~ sum_k w_k N(mean_k, var_k)                                                                  mfs/utils.py:55-58
  double x;
  {
    const int k = mixture_component(draw(traj, 0, 0, P.seed).ua, P.init_weights, P.n_components);
    double z0, z1;
    normals(draw(traj, 0, 1, P.seed), z0, z1);
    x = fma(sqrt(P.init_variances[k]), z0, P.init_means[k]);
  }
  if (P.x0_out) P.x0_out[b] = x;
  YsWriter yw;
  if (P.ys_out) yw.init(P.ys_out, P.ys_dtype, b, P.ys_stride_b, P.ys_stride_t);
  const double c = 0.5 * P.dispersion * P.dispersion;
  const int S = P.integration_steps;
  const double ddt = P.dt / (double)S;
  const double sqdt = sqrt(P.dt);
  for (int64_t t = 0; t < P.T; ++t) {
    const uint32_t ti = (uint32_t)(t + 1);
    if (P.scheme == MFS_SIM_BENES_EXACT) {
      // exact Benes law: X_dt | x ~ N(x + s dt, dt), P(s = +-1) = (1 +- tanh x)/2
      double z0, z1;
      normals(draw(traj, ti, 0, P.seed), z0, z1);
      const double s = draw(traj, ti, kDrawSign, P.seed).ua < 0.5 * (1.0 + tanh(x)) ? 1.0 : -1.0;
      x = fma(sqdt, z0, fma(s, P.dt, x));
    } else {
      // simulate_sde: x <- m(x, ddt) + sqrt(cov(x, ddt)) xi, `integration_steps` times          mfs/utils.py:231-240
      for (int k = 0; k < S; k += 2) {
        double z0, z1, m, v;
        normals(draw(traj, ti, (uint32_t)(k >> 1), P.seed), z0, z1);
        tme_mean_var(drift_jet(P.drift_id, x, tp), x, c, ddt, P.tme_order, m, v);
        x = fma(sqrt(v), z0, m);
        if (k + 1 < S) {
          tme_mean_var(drift_jet(P.drift_id, x, tp), x, c, ddt, P.tme_order, m, v);
          x = fma(sqrt(v), z1, m);
        }
      }
    }
    if (P.xs_out) P.xs_out[b * P.xs_stride_b + t * P.xs_stride_t] = x;
    if (P.ys_out) yw.put(t, P.T, measure(P.meas_id, mp, x, draw(traj, ti, kDrawMeasurement, P.seed)));
  }
}

// Lotka--Volterra, Milstein sub-steps (mfs/multi_dims/ss_models.py:80-86):
//   x <- x + a(x) ddt + sigma x ddw + sigma^2/2 x (ddw^2 - ddt),  a(x) = x * (x[::-1] * [-beta, delta] + [alpha, -gamma])
__global__ void __launch_bounds__(128) simulate_lv_kernel(const mfs_simulate_lv_args P) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const uint64_t traj = (uint64_t)b + P.traj_offset;
  const double* tp = P.trans_params + b * P.trans_param_stride;
  const double alp = tp[0], beta = tp[1], delta = tp[2], gamma = tp[3], sigma = tp[4];
  double mp[2] = {P.meas_params[b * P.meas_param_stride], P.meas_params[b * P.meas_param_stride + 1]};
  double x0, x1;
  {
    // GaussianSumND.sampler: mean_k + chol(cov_k) xi                                              mfs/utils.py:101-105
    const int k = mixture_component(draw(traj, 0, 0, P.seed).ua, P.init_weights, P.n_components);
    double z0, z1;
    normals(draw(traj, 0, 1, P.seed), z0, z1);
    const double l00 = sqrt(P.init_covs[k][0]);
    const double l10 = P.init_covs[k][2] / l00;
    const double l11 = sqrt(P.init_covs[k][3] - l10 * l10);
    x0 = fma(l00, z0, P.init_means[k][0]);
    x1 = fma(l10, z0, fma(l11, z1, P.init_means[k][1]));
  }
  if (P.x0_out) {
    P.x0_out[2 * b] = x0;
    P.x0_out[2 * b + 1] = x1;
  }
  YsWriter yw;
  if (P.ys_out) yw.init(P.ys_out, MFS_YS_U8, b, P.T, 1);
  const int S = P.integration_steps;
  const double ddt = P.dt / (double)S;
  const double sq = sqrt(ddt);
  const double hs2 = 0.5 * sigma * sigma;
  for (int64_t t = 0; t < P.T; ++t) {
    const uint32_t ti = (uint32_t)(t + 1);
    for (int k = 0; k < S; ++k) {
      double z0, z1;
      normals(draw(traj, ti, (uint32_t)k, P.seed), z0, z1);
      const double w0 = sq * z0, w1 = sq * z1;
      const double a0 = x0 * (x1 * -beta + alp);
      const double a1 = x1 * (x0 * delta - gamma);
      const double n0 = x0 + a0 * ddt + sigma * x0 * w0 + hs2 * x0 * (w0 * w0 - ddt);
      const double n1 = x1 + a1 * ddt + sigma * x1 * w1 + hs2 * x1 * (w1 * w1 - ddt);
      x0 = n0;
      x1 = n1;
    }
    if (P.xs_out) {
      P.xs_out[(b * P.T + t) * 2] = x0;
      P.xs_out[(b * P.T + t) * 2 + 1] = x1;
    }
    if (P.ys_out)
      yw.put(t, P.T, measure(MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC, mp, P.obs_dim == 0 ? x0 : x1,
                             draw(traj, ti, kDrawMeasurement, P.seed)));
  }
}

}  // namespace mfs

extern "C" {

int mfs_simulate_1d(const mfs_simulate1d_args* a, void* stream) {
  using namespace mfs;
  if (!a) return fail("mfs_simulate_1d: null argument struct");
  if (a->abi_version != MFS_ABI_VERSION) return fail("mfs_simulate_1d: ABI version %d, library %d", a->abi_version, MFS_ABI_VERSION);
  if (a->B < 0 || a->T < 0) return fail("mfs_simulate_1d: negative B or T");
  if (a->scheme != MFS_SIM_TME && a->scheme != MFS_SIM_BENES_EXACT) return fail("mfs_simulate_1d: unknown scheme %d", a->scheme);
  if (a->scheme == MFS_SIM_TME) {
    if (a->tme_order < 1 || a->tme_order > 3) return fail("mfs_simulate_1d: tme_order must be 1..3, got %d", a->tme_order);
    if (a->integration_steps < 1) return fail("mfs_simulate_1d: integration_steps must be >= 1");
    if (a->drift_id < MFS_DRIFT_BENES || a->drift_id > MFS_DRIFT_LINEAR) return fail("mfs_simulate_1d: unknown drift %d", a->drift_id);
    if (a->drift_id != MFS_DRIFT_BENES && !a->trans_params) return fail("mfs_simulate_1d: this drift needs trans_params");
  }
  if (a->meas_id < MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC || a->meas_id > MFS_MEAS_GAUSSIAN) return fail("mfs_simulate_1d: unknown measurement model %d", a->meas_id);
  if (a->ys_out && !a->meas_params) return fail("mfs_simulate_1d: meas_params is null");
  if (a->ys_dtype < MFS_YS_U8 || a->ys_dtype > MFS_YS_F64) return fail("mfs_simulate_1d: unknown ys dtype %d", a->ys_dtype);
  if (a->meas_id == MFS_MEAS_GAUSSIAN && a->ys_out && a->ys_dtype != MFS_YS_F64) return fail("mfs_simulate_1d: Gaussian measurements need float64 ys");
  if (a->meas_id == MFS_MEAS_POISSON_SOFTPLUS && a->ys_out && a->ys_dtype == MFS_YS_U8) return fail("mfs_simulate_1d: Poisson counts need int32 or float64 ys");
  if (a->n_components < 1 || a->n_components > MFS_SIM_MAX_COMPONENTS) return fail("mfs_simulate_1d: n_components must be 1..%d", MFS_SIM_MAX_COMPONENTS);
  if (!(a->dt > 0.0)) return fail("mfs_simulate_1d: dt must be positive");
  if (a->B == 0) return 0;
  if (!a->ys_out && !a->xs_out && !a->x0_out) return fail("mfs_simulate_1d: no output requested");
  if (a->T == 0 && !a->x0_out) return 0;
  const unsigned grid = (unsigned)((a->B + 127) / 128);
  simulate1d_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("mfs_simulate_1d: launch failed: %s", cudaGetErrorString(e));
  count_launches(1);
  return 0;
}

int mfs_simulate_lv(const mfs_simulate_lv_args* a, void* stream) {
  using namespace mfs;
  if (!a) return fail("mfs_simulate_lv: null argument struct");
  if (a->abi_version != MFS_ABI_VERSION) return fail("mfs_simulate_lv: ABI version %d, library %d", a->abi_version, MFS_ABI_VERSION);
  if (a->B < 0 || a->T < 0) return fail("mfs_simulate_lv: negative B or T");
  if (a->integration_steps < 1) return fail("mfs_simulate_lv: integration_steps must be >= 1");
  if (a->n_components < 1 || a->n_components > MFS_SIM_MAX_COMPONENTS) return fail("mfs_simulate_lv: n_components must be 1..%d", MFS_SIM_MAX_COMPONENTS);
  if (a->obs_dim < 0 || a->obs_dim > 1) return fail("mfs_simulate_lv: obs_dim must be 0 or 1");
  if (!a->trans_params || !a->meas_params) return fail("mfs_simulate_lv: null parameter pointer");
  if (!(a->dt > 0.0)) return fail("mfs_simulate_lv: dt must be positive");
  if (a->B == 0) return 0;
  if (!a->ys_out && !a->xs_out && !a->x0_out) return fail("mfs_simulate_lv: no output requested");
  if (a->T == 0 && !a->x0_out) return 0;
  const unsigned grid = (unsigned)((a->B + 127) / 128);
  simulate_lv_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("mfs_simulate_lv: launch failed: %s", cudaGetErrorString(e));
  count_launches(1);
  return 0;
}

}  // extern "C"
