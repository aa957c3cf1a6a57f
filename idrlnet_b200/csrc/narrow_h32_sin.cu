// Narrow fused-step kernels: hidden width <= 32, activation sin.
#include "narrow_kernel.cuh"
namespace pinn {
PINN_NARROW_TABLE(narrow_table_h32_sin, 32, PINN_ACT_SIN)
}
