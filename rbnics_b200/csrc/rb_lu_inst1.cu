// rb_lu_inst1.cu -- explicit instantiations of the register-resident LU for NP in [56, 60, 64, 68, 72, 76, 80] (split for parallel compilation).
#include "rb_lu_kernel.cuh"

int rb_lu_launch_unit1(int NP, const LuArgs& args, cudaStream_t stream) {
  switch (NP) {
    case 56: return rb_lu_launch_np<56>(args, stream);
    case 60: return rb_lu_launch_np<60>(args, stream);
    case 64: return rb_lu_launch_np<64>(args, stream);
    case 68: return rb_lu_launch_np<68>(args, stream);
    case 72: return rb_lu_launch_np<72>(args, stream);
    case 76: return rb_lu_launch_np<76>(args, stream);
    case 80: return rb_lu_launch_np<80>(args, stream);
    default: return -1000;
  }
}
