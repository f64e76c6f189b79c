// XLA FFI custom-call handlers for the 1-D moment filters: jax.jit callers of the reference keep their code and the scan
// runs in libmfs_b200.so.  One handler per reference entry point (mfs/one_dim/filtering.py:32-36, 92-98, 164-172); each is
// a thin adapter from XLA buffers to mfs_filter1d_args (include/mfs_b200.h), run on XLA's own stream.
//
// Not part of the default build: jax / jaxlib -- which ship xla/ffi/api/ffi.h (jax.ffi.include_dir()) -- are not in this
// image.  With JAX installed:
//
//   g++ -std=c++17 -shared -fPIC -DMFS_WITH_XLA_FFI -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I../../include -I/usr/local/cuda/include ffi_shim.cc -L.. -lmfs_b200 -o ../libmfs_b200_ffi.so
//
// and on the Python side (the stub a maintainer adds next to mfs/one_dim/filtering.py):
//
//   shim = ctypes.CDLL("libmfs_b200_ffi.so")
//   for name in ("MfsFilterRms", "MfsFilterCms", "MfsFilterScms"):
//       jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(shim, name)), platform="CUDA")
//
// Without -DMFS_WITH_XLA_FFI this translation unit is empty (the Makefile compiles it that way so that it stays in the
// source list and a stale include path shows up at once).
#ifdef MFS_WITH_XLA_FFI

#include <cuda_runtime_api.h>

#include "mfs_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// Attributes shared by the three handlers: the named device functors and their scalar constants.
struct Model {
  int32_t trans_id, drift_id, tme_order, meas_id, stable, history;
  double dt, dispersion;
};

// ys: (B, T) uint8 | int32 | float64; ms0: (2N,) or (B, 2N); params: (1, 4) or (B, 4).
template <typename YsBuffer>
ffi::Error Run(cudaStream_t stream, int mode, const Model& m, YsBuffer ys, int32_t ys_dtype, ffi::Buffer<ffi::F64> ms0,
               const double* mean0, int64_t mean0_stride, const double* scale0, int64_t scale0_stride,
               ffi::Buffer<ffi::F64> trans_params, ffi::Buffer<ffi::F64> meas_params, double* ms_out, double* mean_out,
               double* scale_out, double* nell, int32_t* status) {
  if (ys.dimensions().size() != 2) return ffi::Error::InvalidArgument("ys must be (B, T)");
  const int64_t B = ys.dimensions()[0], T = ys.dimensions()[1], M = ms0.dimensions().back();
  mfs_filter1d_args a = {};
  a.abi_version = MFS_ABI_VERSION;
  a.mode = mode;
  a.N = static_cast<int32_t>(M / 2);
  a.stable = m.stable;
  a.B = B;
  a.T = T;
  a.trans_id = m.trans_id;
  a.drift_id = m.drift_id;
  a.tme_order = m.tme_order;
  a.meas_id = m.meas_id;
  a.dt = m.dt;
  a.dispersion = m.dispersion;
  a.trans_params = trans_params.typed_data();
  a.trans_param_stride = trans_params.dimensions()[0] == B && B > 1 ? MFS_MAX_PARAMS : 0;
  a.meas_params = meas_params.typed_data();
  a.meas_param_stride = meas_params.dimensions()[0] == B && B > 1 ? MFS_MAX_PARAMS : 0;
  a.ms0 = ms0.typed_data();
  a.ms0_stride = ms0.dimensions().size() == 2 && ms0.dimensions()[0] == B && B > 1 ? M : 0;
  a.mean0 = mean0;
  a.mean0_stride = mean0_stride;
  a.scale0 = scale0;
  a.scale0_stride = scale0_stride;
  a.ys = ys.untyped_data();
  a.ys_dtype = ys_dtype;
  a.ys_stride_b = T;
  a.ys_stride_t = 1;
  a.out_mode = m.history;                       // MFS_OUT_FULL (the reference's return value) | LAST | NONE | MEANVAR
  a.ms_out = ms_out;
  if (m.history == MFS_OUT_FULL) { a.ms_stride_b = T * M; a.ms_stride_t = M; a.aux_stride_b = T; }
  if (m.history == MFS_OUT_MEANVAR) { a.ms_stride_b = T * 2; a.ms_stride_t = 2; a.aux_stride_b = T; }
  if (m.history == MFS_OUT_LAST) { a.ms_stride_b = M; a.aux_stride_b = 1; }
  a.mean_out = mean_out;
  a.scale_out = scale_out;
  a.nell_out = nell;
  a.status_out = status;
  if (mfs_filter_1d(&a, stream) != 0) return ffi::Error::Internal(mfs_last_error());
  return ffi::Error::Success();
}

int64_t PerFilterStride(const ffi::Buffer<ffi::F64>& v, int64_t B) { return v.element_count() == B && B > 1 ? 1 : 0; }

// moment_filter_rms(state_cond_raw_moments, measurement_cond_pdf, rms0, ys) -> (rmss, nell)   filtering.py:32-89
ffi::Error FilterRmsImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> ys, ffi::Buffer<ffi::F64> rms0,
                         ffi::Buffer<ffi::F64> trans_params, ffi::Buffer<ffi::F64> meas_params, int32_t trans_id,
                         int32_t drift_id, int32_t tme_order, int32_t meas_id, int32_t stable, int32_t history, double dt,
                         double dispersion, ffi::ResultBuffer<ffi::F64> rmss, ffi::ResultBuffer<ffi::F64> nell,
                         ffi::ResultBuffer<ffi::S32> status) {
  const Model m{trans_id, drift_id, tme_order, meas_id, stable, history, dt, dispersion};
  return Run(stream, MFS_MODE_RAW, m, ys, MFS_YS_U8, rms0, nullptr, 0, nullptr, 0, trans_params, meas_params,
             rmss->typed_data(), nullptr, nullptr, nell->typed_data(), status->typed_data());
}

// moment_filter_cms(..., cms0, mean0, ys) -> (cmss, means, nell)                              filtering.py:92-161
ffi::Error FilterCmsImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> ys, ffi::Buffer<ffi::F64> cms0,
                         ffi::Buffer<ffi::F64> mean0, ffi::Buffer<ffi::F64> trans_params,
                         ffi::Buffer<ffi::F64> meas_params, int32_t trans_id, int32_t drift_id, int32_t tme_order,
                         int32_t meas_id, int32_t stable, int32_t history, double dt, double dispersion,
                         ffi::ResultBuffer<ffi::F64> cmss, ffi::ResultBuffer<ffi::F64> means,
                         ffi::ResultBuffer<ffi::F64> nell, ffi::ResultBuffer<ffi::S32> status) {
  const Model m{trans_id, drift_id, tme_order, meas_id, stable, history, dt, dispersion};
  const int64_t B = ys.dimensions()[0];
  return Run(stream, MFS_MODE_CENTRAL, m, ys, MFS_YS_U8, cms0, mean0.typed_data(), PerFilterStride(mean0, B), nullptr, 0,
             trans_params, meas_params, cmss->typed_data(), means->typed_data(), nullptr, nell->typed_data(),
             status->typed_data());
}

// moment_filter_scms(..., scms0, mean0, scale0, ys) -> (scmss, means, scales, nell)           filtering.py:164-240
ffi::Error FilterScmsImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> ys, ffi::Buffer<ffi::F64> scms0,
                          ffi::Buffer<ffi::F64> mean0, ffi::Buffer<ffi::F64> scale0,
                          ffi::Buffer<ffi::F64> trans_params, ffi::Buffer<ffi::F64> meas_params, int32_t trans_id,
                          int32_t drift_id, int32_t tme_order, int32_t meas_id, int32_t stable, int32_t history,
                          double dt, double dispersion, ffi::ResultBuffer<ffi::F64> scmss,
                          ffi::ResultBuffer<ffi::F64> means, ffi::ResultBuffer<ffi::F64> scales,
                          ffi::ResultBuffer<ffi::F64> nell, ffi::ResultBuffer<ffi::S32> status) {
  const Model m{trans_id, drift_id, tme_order, meas_id, stable, history, dt, dispersion};
  const int64_t B = ys.dimensions()[0];
  return Run(stream, MFS_MODE_SCALED, m, ys, MFS_YS_U8, scms0, mean0.typed_data(), PerFilterStride(mean0, B),
             scale0.typed_data(), PerFilterStride(scale0, B), trans_params, meas_params, scmss->typed_data(),
             means->typed_data(), scales->typed_data(), nell->typed_data(), status->typed_data());
}

#define MFS_MODEL_ATTRS()                                                                                      \
  Attr<int32_t>("trans_id").Attr<int32_t>("drift_id").Attr<int32_t>("tme_order").Attr<int32_t>("meas_id")     \
      .Attr<int32_t>("stable").Attr<int32_t>("history").Attr<double>("dt").Attr<double>("dispersion")

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(MfsFilterRms, FilterRmsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()    // ys
                                  .Arg<ffi::Buffer<ffi::F64>>()   // rms0
                                  .Arg<ffi::Buffer<ffi::F64>>()   // trans_params
                                  .Arg<ffi::Buffer<ffi::F64>>()   // meas_params
                                  .MFS_MODEL_ATTRS()
                                  .Ret<ffi::Buffer<ffi::F64>>()   // rmss
                                  .Ret<ffi::Buffer<ffi::F64>>()   // nell
                                  .Ret<ffi::Buffer<ffi::S32>>()); // status

XLA_FFI_DEFINE_HANDLER_SYMBOL(MfsFilterCms, FilterCmsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // cms0
                                  .Arg<ffi::Buffer<ffi::F64>>()   // mean0
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .MFS_MODEL_ATTRS()
                                  .Ret<ffi::Buffer<ffi::F64>>()   // cmss
                                  .Ret<ffi::Buffer<ffi::F64>>()   // means
                                  .Ret<ffi::Buffer<ffi::F64>>()   // nell
                                  .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(MfsFilterScms, FilterScmsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // scms0
                                  .Arg<ffi::Buffer<ffi::F64>>()   // mean0
                                  .Arg<ffi::Buffer<ffi::F64>>()   // scale0
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .MFS_MODEL_ATTRS()
                                  .Ret<ffi::Buffer<ffi::F64>>()   // scmss
                                  .Ret<ffi::Buffer<ffi::F64>>()   // means
                                  .Ret<ffi::Buffer<ffi::F64>>()   // scales
                                  .Ret<ffi::Buffer<ffi::F64>>()   // nell
                                  .Ret<ffi::Buffer<ffi::S32>>());

#endif  // MFS_WITH_XLA_FFI
