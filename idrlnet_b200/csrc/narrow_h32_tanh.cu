// Narrow fused-step kernels: hidden width <= 32, activation tanh.
#include "narrow_kernel.cuh"
namespace pinn {
PINN_NARROW_TABLE(narrow_table_h32_tanh, 32, PINN_ACT_TANH)
}
