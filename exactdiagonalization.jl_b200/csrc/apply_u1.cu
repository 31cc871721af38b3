// K2 fast path (spin-1/2, single U(1) sector, combinadic ranking) -- placeholder until the tiled kernel lands.
#include "ed_device.cuh"

bool ed_apply_u1_supported(ed_oprep* o, int dtype, int side) { (void)o; (void)dtype; (void)side; return false; }
void ed_apply_u1(ed_oprep* o, void* out, const void* x, int dtype, int side, int accumulate, double* alpha_dot) {
  (void)o; (void)out; (void)x; (void)dtype; (void)side; (void)accumulate; (void)alpha_dot;
  throw EdError(ED_ERR_INTERNAL, "u1 fast path not built");
}
