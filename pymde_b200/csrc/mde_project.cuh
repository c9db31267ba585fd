// mde_project.cuh -- constraint projections shared by the C ABI (mde_project.cu) and the
// solver (mde_solver.cu).  Reference: pymde/constraints.py:94-200, pymde/util.py:129-171.
#pragma once
#include "mde_common.cuh"

namespace mde {

constexpr int kProjMaxM = 32;    // Standardized on device: m <= 32 (Gram + warp Jacobi)
constexpr int kProjBlocks = kNumSMs * 2;
constexpr int kProjThreads = 256;

// workspace (doubles): [0, kProjBlocks * K) block partials, then finals
struct ProjWs {
  double* partials;  // kProjBlocks * kmax
  double* mean;      // m            (column means)
  double* mat;       // m*m          (W for the retraction, or Z^T X / n for the tangent)
  int* status;       // 1 int: 0 ok, 1 = Gram not positive definite
};

inline int64_t proj_mm(int m) { return m <= kProjMaxM ? (int64_t)m * m : 0; }

inline int64_t proj_ws_doubles(int m) {
  int64_t k = (int64_t)m + proj_mm(m);
  return (int64_t)kProjBlocks * k + m + proj_mm(m) + 8;
}

inline ProjWs proj_ws_carve(void* ws, int m) {
  ProjWs w;
  int64_t k = (int64_t)m + proj_mm(m);
  w.partials = (double*)ws;
  w.mean = w.partials + (int64_t)kProjBlocks * k;
  w.mat = w.mean + m;
  w.status = (int*)(w.mat + proj_mm(m));
  return w;
}

// Enqueue X -= colmean(X).  `active` (nullable) is a device flag; kernels exit when it is 0.
int enqueue_project_centered(float* X, int64_t n, int m, const ProjWs& w, const int* active, cudaStream_t st);
// Enqueue de-mean + sqrt(n) * polar factor.  m <= kProjMaxM.
int enqueue_project_standardized(float* X, int64_t n, int m, const ProjWs& w, const int* active, cudaStream_t st);
// Enqueue Z -= (1/n) X (Z^T X).  m <= kProjMaxM.
int enqueue_tangent_standardized(const float* X, float* Z, int64_t n, int m, const ProjWs& w,
                                 const int* active, cudaStream_t st);

}  // namespace mde
