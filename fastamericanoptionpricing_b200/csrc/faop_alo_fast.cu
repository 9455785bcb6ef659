// placeholder, replaced by the specialised kernels
#include "faop_kernels.h"
namespace faop {
int launch_alo_fast(const AloArgs& a, cudaStream_t st) { (void)a; (void)st; return -1000; }
}
