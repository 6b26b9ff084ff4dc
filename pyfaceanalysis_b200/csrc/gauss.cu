// Gaussian-classifier regression / label head (sm_100a, float64).
//
// Replaces classifiers[i].regression(sl[:, 0:D], avg_labels[, estimate_std]) and .label(x)
// (reference FaceDetectUpdated.py:709-719, face_analysis.py:1068-1071,1261-1287), i.e.
// mdp.nodes.GaussianClassifier.class_probabilities + cuicuilco's GaussianRegression
// (SURVEY.md row a-12):
//     q_c   = p_c * (2 pi)^(-D/2) / sqrt_det_c * exp(-1/2 (x - mu_c)^T S_c^-1 (x - mu_c))
//     P_c   = q_c / sum_c q_c
//     value = sum_c P_c l_c ;  std = sqrt(sum_c P_c (l_c - value)^2) ;  winner = argmax_c P_c
// Evaluated in float64 in that order (not in the log domain) so that the reference's underflow
// behaviour -- every q_c == 0 -> 0/0 -> NaN -> "NaN >= cut_off" is False -> window kept -- is
// reproduced rather than repaired.
//
// One thread per window; (x - mu_c) lives in registers (D is a template bucket), S_c^-1 is staged in
// shared memory class group by class group and read as warp-uniform broadcasts; q_c of all classes
// is parked in shared memory ([C][128 threads]) until the normalising sum is known.
#include <vector>

#include "common.cuh"

namespace hgsfa {

constexpr int GT = 128;  // threads (= windows) per CTA

template <int DMAX, typename XT>
__global__ void __launch_bounds__(GT) gauss_kernel(const XT* __restrict__ x, int64_t n, int64_t ld, int D, int C,
                                                   int group, const double* __restrict__ means,
                                                   const double* __restrict__ inv_covs,
                                                   const double* __restrict__ consts,  // p_c-free constant
                                                   const double* __restrict__ priors,
                                                   const double* __restrict__ avg_labels, double* __restrict__ value,
                                                   double* __restrict__ stdv, int* __restrict__ winner,
                                                   double* __restrict__ probs) {
  extern __shared__ __align__(16) double sm[];
  double* sS = sm;                          // [group][D][D]
  double* sMu = sS + size_t(group) * D * D; // [group][D]
  double* sQ = sMu + size_t(group) * D;     // [C][GT]
  const int tid = threadIdx.x;
  const int64_t w = int64_t(blockIdx.x) * GT + tid;
  const bool live = w < n;

  double xv[DMAX];
#pragma unroll
  for (int j = 0; j < DMAX; ++j) xv[j] = (live && j < D) ? double(x[w * ld + j]) : 0.0;

  for (int c0 = 0; c0 < C; c0 += group) {
    const int gc = min(group, C - c0);
    __syncthreads();
    for (int i = tid; i < gc * D * D; i += GT) sS[i] = inv_covs[size_t(c0) * D * D + i];
    for (int i = tid; i < gc * D; i += GT) sMu[i] = means[size_t(c0) * D + i];
    __syncthreads();
    for (int c = 0; c < gc; ++c) {
      double d[DMAX];
#pragma unroll
      for (int j = 0; j < DMAX; ++j) d[j] = (j < D) ? xv[j] - sMu[c * D + j] : 0.0;
      // exponent = 0.5 * sum_i (sum_j d_j S[j][i]) d_i      (x_mn @ invS, then * x_mn, summed)
      double e = 0.0;
      const double* S = sS + size_t(c) * D * D;
#pragma unroll
      for (int i = 0; i < DMAX; ++i) {
        if (i < D) {
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < DMAX; ++j)
            if (j < D) t = fma(d[j], S[j * D + i], t);
          e = fma(t, d[i], e);
        }
      }
      e *= 0.5;
      double q = consts[c0 + c] * exp(-e);
      q *= priors[c0 + c];
      sQ[(c0 + c) * GT + tid] = q;
    }
  }
  if (!live) return;
  double tot = 0.0;
  for (int c = 0; c < C; ++c) tot += sQ[c * GT + tid];
  double val = 0.0, best = -1.0;
  int arg = 0;
  for (int c = 0; c < C; ++c) {
    const double P = sQ[c * GT + tid] / tot;   // 0/0 -> NaN exactly like the reference
    if (probs) probs[w * C + c] = P;
    if (avg_labels) val = fma(P, avg_labels[c], val);
    if (P > best) { best = P; arg = c; }       // first maximum, like numpy argmax (NaN never wins)
  }
  if (value) value[w] = val;
  if (winner) winner[w] = arg;
  if (stdv) {
    double s = 0.0;
    for (int c = 0; c < C; ++c) {
      const double P = sQ[c * GT + tid] / tot;
      const double df = avg_labels[c] - val;
      s = fma(P, df * df, s);
    }
    stdv[w] = sqrt(s);
  }
}

}  // namespace hgsfa

using namespace hgsfa;

struct hgsfa_gauss_s {
  int device = 0, C = 0, D = 0;
  cudaStream_t stream = nullptr;
  DevBuf means, inv_covs, consts, priors, labels;
  DevBuf sx, sval, sstd, swin, sprob;  // staging for the host entry point
  int64_t launches = 0;
};

extern "C" int hgsfa_gauss_create(const double* means, const double* inv_covs, const double* sqrt_det,
                                  const double* priors, int C, int D, int device, hgsfa_gauss_t* out) {
  HG_CHECK(means && inv_covs && sqrt_det && priors && out, "hgsfa_gauss_create: null argument");
  HG_CHECK(C >= 1 && C <= 4096, "hgsfa_gauss_create: class count %d out of range", C);
  HG_CHECK(D >= 1 && D <= 32, "hgsfa_gauss_create: input_dim %d out of range [1, 32]", D);
  int ndev = 0;
  HG_CUDA(cudaGetDeviceCount(&ndev));
  HG_CHECK(device >= 0 && device < ndev, "hgsfa_gauss_create: device %d out of range (%d devices)", device, ndev);
  DeviceGuard guard(device);
  auto h = new hgsfa_gauss_s();
  h->device = device; h->C = C; h->D = D;
  // constant = (2 pi)^(-D/2) / sqrt_det_c, evaluated in double on the host exactly as numpy does
  std::vector<double> consts(C);
  const double base = pow(2.0 * 3.141592653589793, -double(D) / 2.0);
  for (int c = 0; c < C; ++c) consts[c] = base / sqrt_det[c];
  bool ok = !h->means.reserve(sizeof(double) * C * D) && !h->inv_covs.reserve(sizeof(double) * C * D * D) &&
            !h->consts.reserve(sizeof(double) * C) && !h->priors.reserve(sizeof(double) * C) &&
            !h->labels.reserve(sizeof(double) * C);
  ok = ok && cudaMemcpy(h->means.p, means, sizeof(double) * C * D, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(h->inv_covs.p, inv_covs, sizeof(double) * C * D * D, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(h->consts.p, consts.data(), sizeof(double) * C, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(h->priors.p, priors, sizeof(double) * C, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  if (!ok) {
    hgsfa_gauss_destroy(h);
    return fail("hgsfa_gauss_create: device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = h;
  return 0;
}

extern "C" int hgsfa_gauss_destroy(hgsfa_gauss_t h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  h->means.release(); h->inv_covs.release(); h->consts.release(); h->priors.release(); h->labels.release();
  h->sx.release(); h->sval.release(); h->sstd.release(); h->swin.release(); h->sprob.release();
  delete h;
  return 0;
}

namespace {

template <typename XT>
int launch_gauss(hgsfa_gauss_t h, const XT* x, int64_t n, int64_t ld, const double* labels, double* value,
                 double* stdv, int* winner, double* probs, cudaStream_t st) {
  const int C = h->C, D = h->D;
  // class group: as many inverse covariances as fit next to the q table
  const size_t q_bytes = sizeof(double) * C * GT;
  const size_t budget = 200 * 1024;
  HG_CHECK(q_bytes + sizeof(double) * (D * D + D) <= budget, "hgsfa_gauss: %d classes need too much shared memory", C);
  int group = int((budget - q_bytes) / (sizeof(double) * (D * D + D)));
  if (group > C) group = C;
  const size_t smem = q_bytes + sizeof(double) * size_t(group) * (D * D + D);
  const unsigned grid = (unsigned)ceil_div(n, GT);
#define HG_GAUSS(DM)                                                                                            \
  do {                                                                                                          \
    /* the limit is per function and device, not per launch: always the full budget, so that concurrent callers with    \
       classifiers of different sizes (one host thread per stream) cannot lower it under each other's launches */       \
    HG_CUDA(cudaFuncSetAttribute(gauss_kernel<DM, XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)); \
    gauss_kernel<DM, XT><<<grid, GT, smem, st>>>(x, n, ld, D, C, group, (const double*)h->means.p,             \
                                                 (const double*)h->inv_covs.p, (const double*)h->consts.p,     \
                                                 (const double*)h->priors.p, labels, value, stdv, winner, probs); \
  } while (0)
  if (D <= 8) HG_GAUSS(8);
  else if (D <= 12) HG_GAUSS(12);
  else if (D <= 16) HG_GAUSS(16);
  else if (D <= 20) HG_GAUSS(20);
  else if (D <= 24) HG_GAUSS(24);
  else HG_GAUSS(32);
#undef HG_GAUSS
  h->launches++;
  HG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int hgsfa_gauss_regress_device(hgsfa_gauss_t h, const void* d_x, int x_dtype, int64_t n, int64_t ld,
                                          const double* d_avg_labels, double* d_value, double* d_std,
                                          int32_t* d_winner, double* d_probs, void* stream) {
  HG_CHECK(h, "hgsfa_gauss_regress: null handle");
  HG_CHECK(n >= 0, "hgsfa_gauss_regress: negative row count");
  HG_CHECK(ld >= h->D, "GaussianClassifier: x has dimension %lld, should be %d", (long long)ld, h->D);
  HG_CHECK(x_dtype == HGSFA_F32 || x_dtype == HGSFA_F64, "hgsfa_gauss_regress: x dtype must be f32 or f64");
  HG_CHECK(!(d_value || d_std) || d_avg_labels, "hgsfa_gauss_regress: value / std requested without avg_labels");
  if (n == 0) return 0;
  HG_CHECK(d_x, "hgsfa_gauss_regress: null x");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);  // NULL = default stream (CUDA semantics)
  if (x_dtype == HGSFA_F32)
    return launch_gauss<float>(h, static_cast<const float*>(d_x), n, ld, d_avg_labels, d_value, d_std, d_winner, d_probs, st);
  return launch_gauss<double>(h, static_cast<const double*>(d_x), n, ld, d_avg_labels, d_value, d_std, d_winner, d_probs, st);
}

extern "C" int hgsfa_gauss_regress(hgsfa_gauss_t h, const void* x, int x_dtype, int64_t n, int64_t ld,
                                   const double* avg_labels, double* value, double* stdv, int32_t* winner,
                                   double* probs, void* stream) {
  HG_CHECK(h, "hgsfa_gauss_regress: null handle");
  HG_CHECK(n >= 0, "hgsfa_gauss_regress: negative row count");
  HG_CHECK(ld >= h->D, "GaussianClassifier: x has dimension %lld, should be %d", (long long)ld, h->D);
  HG_CHECK(x_dtype == HGSFA_F32 || x_dtype == HGSFA_F64, "hgsfa_gauss_regress: x dtype must be f32 or f64");
  HG_CHECK(!(value || stdv) || avg_labels, "hgsfa_gauss_regress: value / std requested without avg_labels");
  if (n == 0) return 0;
  HG_CHECK(x, "hgsfa_gauss_regress: null x");
  DeviceGuard guard(h->device);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  const size_t xel = dtype_size(x_dtype);
  const int C = h->C, D = h->D;
  if (h->sx.reserve(size_t(n) * D * xel)) return 1;
  if (value && h->sval.reserve(size_t(n) * 8)) return 1;
  if (stdv && h->sstd.reserve(size_t(n) * 8)) return 1;
  if (winner && h->swin.reserve(size_t(n) * 4)) return 1;
  if (probs && h->sprob.reserve(size_t(n) * C * 8)) return 1;
  HG_CUDA(cudaMemcpy2DAsync(h->sx.p, size_t(D) * xel, x, size_t(ld) * xel, size_t(D) * xel, size_t(n),
                            cudaMemcpyHostToDevice, st));
  if (avg_labels) HG_CUDA(cudaMemcpyAsync(h->labels.p, avg_labels, sizeof(double) * C, cudaMemcpyHostToDevice, st));
  if (hgsfa_gauss_regress_device(h, h->sx.p, x_dtype, n, D, avg_labels ? (const double*)h->labels.p : nullptr,
                                 value ? (double*)h->sval.p : nullptr, stdv ? (double*)h->sstd.p : nullptr,
                                 winner ? (int32_t*)h->swin.p : nullptr, probs ? (double*)h->sprob.p : nullptr, st))
    return 1;
  if (value) HG_CUDA(cudaMemcpyAsync(value, h->sval.p, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
  if (stdv) HG_CUDA(cudaMemcpyAsync(stdv, h->sstd.p, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
  if (winner) HG_CUDA(cudaMemcpyAsync(winner, h->swin.p, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
  if (probs) HG_CUDA(cudaMemcpyAsync(probs, h->sprob.p, size_t(n) * C * 8, cudaMemcpyDeviceToHost, st));
  HG_CUDA(cudaStreamSynchronize(st));
  return 0;
}
