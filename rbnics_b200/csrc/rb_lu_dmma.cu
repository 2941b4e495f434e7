// rb_lu_dmma.cu -- instantiations of the DMMA-blocked LU (rb_lu_dmma.cuh) for 5 <= NT <= 13 tile rows (32 < N <= 104).
#include "rb_lu_dmma.cuh"

#define LUD_CASE(NT_)                                                        \
  case NT_:                                                                  \
    return (NC == NT_) ? rb_lu_dmma_launch<NT_, NT_>(args, stream)           \
                       : rb_lu_dmma_launch<NT_, NT_ + 1>(args, stream);

int rb_lu_dmma_dispatch(int N, const LuArgs& args, cudaStream_t stream) {
  const int NT = (N + 7) / 8, NC = (N + 8) / 8;
  switch (NT) {
    LUD_CASE(5)
    LUD_CASE(6)
    LUD_CASE(7)
    LUD_CASE(8)
    LUD_CASE(9)
    LUD_CASE(10)
    LUD_CASE(11)
    LUD_CASE(12)
    LUD_CASE(13)
    default:
      return -1000;
  }
}
