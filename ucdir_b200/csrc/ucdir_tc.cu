// tcgen05 / TMA bf16 kernels (placeholder until the tensor-core path lands; ops report "unsupported").
#include "common.cuh"
namespace ucdir {
int launch_tc_conv(const ucdir_op_t&, cudaStream_t, bool) { set_error("TC_CONV not built yet"); return -2; }
int launch_tc_attn(const ucdir_op_t&, cudaStream_t, bool) { set_error("TC_ATTN not built yet"); return -2; }
int launch_gn_apply_bf16(const ucdir_op_t&, cudaStream_t, bool) { set_error("GN_APPLY_BF16 not built yet"); return -2; }
int launch_cast(const ucdir_op_t&, cudaStream_t, bool) { set_error("CAST not built yet"); return -2; }
}
