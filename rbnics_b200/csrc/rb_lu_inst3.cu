// rb_lu_inst3.cu -- explicit instantiations of the register-resident LU for NP in [100, 104] (split for parallel compilation).
#include "rb_lu_kernel.cuh"

int rb_lu_launch_unit3(int NP, const LuArgs& args, cudaStream_t stream) {
  switch (NP) {
    case 100: return rb_lu_launch_np<100>(args, stream);
    case 104: return rb_lu_launch_np<104>(args, stream);
    default: return -1000;
  }
}
