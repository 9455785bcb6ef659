// faop_cn.cu — batched Crank-Nicolson American put solver (CNOptionSolver, CrankNicolsonOptionSolver.py:5-107).
#include "faop_kernels.h"
extern "C" {
int faop_cn_workspace_bytes(int64_t N, int32_t n_space, int32_t mode, size_t* bytes) { if (bytes) *bytes = 0; return 0; }
int faop_cn_put_batch(int64_t N, int32_t n_space, int32_t mode, const double* r, const double* q, const double* sigma,
                      const double* K, const double* S0, const double* dts, const int32_t* n_steps, int32_t max_steps,
                      double omega, double tol, int32_t max_iter, double* price, double* X_out, void* workspace,
                      void* cuda_stream) { return FAOP_ERANGE; }
}
