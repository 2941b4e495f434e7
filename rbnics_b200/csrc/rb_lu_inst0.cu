// rb_lu_inst0.cu -- explicit instantiations of the register-resident LU for NP in [4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52] (split for parallel compilation).
#include "rb_lu_kernel.cuh"

int rb_lu_launch_unit0(int NP, const LuArgs& args, cudaStream_t stream) {
  switch (NP) {
    case 4: return rb_lu_launch_np<4>(args, stream);
    case 8: return rb_lu_launch_np<8>(args, stream);
    case 12: return rb_lu_launch_np<12>(args, stream);
    case 16: return rb_lu_launch_np<16>(args, stream);
    case 20: return rb_lu_launch_np<20>(args, stream);
    case 24: return rb_lu_launch_np<24>(args, stream);
    case 28: return rb_lu_launch_np<28>(args, stream);
    case 32: return rb_lu_launch_np<32>(args, stream);
#ifdef RB_AB_VARIANTS
    case 36: return rb_lu_launch_np<36>(args, stream);
    case 40: return rb_lu_launch_np<40>(args, stream);
    case 44: return rb_lu_launch_np<44>(args, stream);
    case 48: return rb_lu_launch_np<48>(args, stream);
    case 52: return rb_lu_launch_np<52>(args, stream);
#endif
    default: return -1000;
  }
}
