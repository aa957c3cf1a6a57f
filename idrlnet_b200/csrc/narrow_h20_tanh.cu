// Narrow fused-step kernels: hidden width <= 20, activation tanh.
#include "narrow_kernel.cuh"
namespace pinn {
PINN_NARROW_TABLE(narrow_table_h20_tanh, 20, PINN_ACT_TANH)
}
