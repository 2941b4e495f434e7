// rb_lu_inst2.cu -- explicit instantiations of the register-resident LU for NP in [84, 88, 92, 96] (split for parallel compilation).
#include "rb_lu_kernel.cuh"

int rb_lu_launch_unit2(int NP, const LuArgs& args, cudaStream_t stream) {
  switch (NP) {
    case 84: return rb_lu_launch_np<84>(args, stream);
    case 88: return rb_lu_launch_np<88>(args, stream);
    case 92: return rb_lu_launch_np<92>(args, stream);
    case 96: return rb_lu_launch_np<96>(args, stream);
    default: return -1000;
  }
}
