// Single translation unit -> one sm_100a cubin, embedded into libaule.so and loaded
// with cuModuleLoadData (the analogue of the reference's @embedFile'd SPIR-V,
// src/lib.zig:29-50, and of hipModuleLoadData in src/backends/hip.zig:133-160).
#include "attn_simt.cu"
#include "attn_fwd_sm100.cu"
#include "attn_fwd_tf32_sm100.cu"
#ifdef AULE_TUNING_VARIANTS
#include "attn_fwd_sm100_v4.cu"
#endif
#include "attn_bwd_sm100.cu"
#include "attn_bwd_fused_sm100.cu"
#include "attn_bwd_fused2_sm100.cu"
#include "attn_paged_sm100.cu"
