// Kernel parameter blocks and shared-memory layouts shared by the device code
// (csrc/kernels/*.cu) and the host launcher (csrc/host/engine.cpp).  Plain structs only.
#pragma once
#include <stdint.h>

namespace aule_kp {

// ---- CUDA-core kernels (attn_simt.cu) -------------------------------------------------
struct SimtParams {
    const void* q; const void* k; const void* v; void* o;
    float* lse;                  // [B,Hq,Sq] or nullptr (forward) / required (backward)
    const void* d_o; void* dq; void* dk; void* dv;
    float* delta;                // [B,Hq,Sq] workspace (backward)
    uint32_t B, Hq, Hkv, Sq, Sk, D;
    float scale;
    int32_t causal, window;
};
constexpr int SIMT_ROWS = 32, SIMT_LANES = 8;

// ---- RoPE of Q and K in one launch (attn_simt.cu: aule_rope_*) -------------------------
struct RopeParams {
    const void* xq; void* oq; uint64_t rows_q; uint32_t Sq;     // [rows_q = B*Hq*Sq, D]
    const void* xk; void* ok; uint64_t rows_k; uint32_t Sk;     // second tensor (rows_k = 0: none)
    const float* cs; const float* sn;                           // [table_rows >= max(Sq, Sk), D/2] fp32
    uint32_t D; int32_t mode;                                   // 0 half-split (Triton path), 1 interleaved pairs (Vulkan shader)
    float sign;                                                 // -1: inverse rotation (backward)
};

// ---- tcgen05 forward (attn_fwd_sm100.cu) ----------------------------------------------
struct FwdParams {
    float* lse;               // [B,Hq,Sq] fp32 or nullptr
    void* o;                  // [B,Hq,Sq,D_real] output (v5 stores it from registers)
    uint32_t B, Hq, Hkv, Sq, Sk;
    uint32_t D_real;          // head_dim of the tensors in memory (<= the kernel's padded D; the TMA boxes zero-fill the rest)
    uint32_t num_q_super;     // ceil(Sq / 256)
    uint32_t num_tiles;       // num_q_super * Hq * B
    float scale;              // softmax scale (natural)
    float scale_log2;         // scale * log2(e)
    // Visibility: key j contributes to query i iff  i - win_left <= j <= i + win_right  (and j < Sk).
    //   causal: win_right = 0; causal window W: win_left = W-1; bidirectional window W: win_left = win_right = W/2
    //   (attention_f32.comp:173-183); "unbounded" = kWinInf.
    uint32_t win_left, win_right;
    int32_t causal;           // v4 kernel only
    int32_t window;           // v4 kernel only
    int32_t pair_heads;       // 1: a work item is 128 rows x 2 adjacent q-heads of one KV group (equal trip counts);
                              // 0: 256 rows of one q-head
    uint32_t units_per_run;   // (batch, kv-head) units whose work items are scheduled together (L2 residency of K/V)
    uint32_t* sched_counter;  // zero-initialised per launch: next unclaimed work item (dynamic persistent scheduler)
    int32_t cross_item;       // v4 only: the first Q K^T of the next work item is issued under the current item's last block
    unsigned long long* trace; // bring-up: CTA 0 records (tag << 48 | clock64) events here (4 x 4096 entries) or nullptr
};
constexpr uint32_t kWinInf = 0x40000000u;

// v5 layout: Q ring (3 tiles) | K/V ring (NS x [128 keys][D], K and V tiles interleaved, continuous across work
// items) | row statistics | work descriptors (8 x 32 B) | mbarriers.  TMEM: one shared S buffer, P0, P1, O0, O1.
template <int D>
struct FwdCfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128 on the tensor-core path");
    static constexpr int NS = (D == 128) ? 4 : 8;               // K/V ring stages
    static constexpr int NQ = 3;                                // Q ring slots: tile x of item k -> slot (2k + x) % 3
    static constexpr int CHUNKS = D / 64;                       // 128-byte swizzle chunks per row
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;          // [128 rows][128 B]
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_Q = 0;
    static constexpr uint32_t OFF_KV = OFF_Q + NQ * TILE_BYTES;
    // D = 64 only: the softmax row sums come out of the P V MMA itself.  A constant [128 keys][64] tile whose column 0 is 1.0
    // sits behind the K/V ring; the V descriptor's leading-dimension offset points the MMA's columns [64,80) at it, so
    // O gets a 65th column = sum_j P_ij -- the 64 FADD2 per row and block leave the (issue-bound) softmax warps, the
    // tensor pipe (half idle at D = 64) pays 25 % more P V work.  D = 128 has no TMEM columns left for it.
    static constexpr bool ROWSUM_MMA = (D == 64);
    static constexpr uint32_t OFF_ONES = OFF_KV + NS * TILE_BYTES;
    static constexpr uint32_t OFF_STAT = OFF_ONES + (ROWSUM_MMA ? CHUNK_BYTES : 0); // float l[2][128], m[2][128]
    static constexpr uint32_t OFF_WORK = OFF_STAT + 4 * 128 * 4;   // 8 work descriptors x 32 B
    static constexpr uint32_t OFF_BAR = OFF_WORK + 8 * 32;
    static constexpr int NBAR = 39 + 2 * NS;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + NBAR * 8;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
    // TMEM columns (512 allocated)
    static constexpr uint32_t COL_S = 0, COL_P0 = 128, COL_P1 = 192, COL_O0 = 256, COL_O1 = ROWSUM_MMA ? 384 : 256 + D;
    static constexpr uint32_t PV_N = ROWSUM_MMA ? D + 16 : D;     // columns of the P V MMA (row sum in column D)
    // register budget per thread after setmaxnreg (512 threads x 128 at launch = 64 K registers):
    // 2 softmax warpgroups x 184 + epilogue warpgroup x 64 + issuer/producer warpgroup x 80 = 512 x 128
    static constexpr int REGS_SOFTMAX = 184, REGS_EPILOGUE = 64, REGS_OTHER = 80;
};

// tf32 forward (attn_fwd_tf32_sm100.cu): fp32 operands, head_dim <= 64.  An fp32 [128][64] tile has the bytes of a bf16
// [128][128] one: shared-memory geometry of FwdCfg<128>; P stays fp32 in TMEM (128 columns per tile).
struct FwdCfgT32 {
    static constexpr int D = 64;                                // logical (padded) head_dim
    static constexpr int NS = 4, NQ = 3;
    static constexpr int CHUNKS = 2;                            // 128-byte swizzle chunks per row: 2 x 32 fp32
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_Q = 0;
    static constexpr uint32_t OFF_KV = OFF_Q + NQ * TILE_BYTES;
    static constexpr bool ROWSUM_MMA = false;
    static constexpr uint32_t OFF_ONES = OFF_KV + NS * TILE_BYTES;
    static constexpr uint32_t OFF_STAT = OFF_ONES;
    static constexpr uint32_t OFF_WORK = OFF_STAT + 4 * 128 * 4;
    static constexpr uint32_t OFF_BAR = OFF_WORK + 8 * 32;
    static constexpr int NBAR = 39 + 2 * NS;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + NBAR * 8;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
    static constexpr uint32_t COL_S = 0, COL_P0 = 128, COL_P1 = 256, COL_O0 = 384, COL_O1 = 448;
    static constexpr uint32_t PV_N = D;
    static constexpr int REGS_SOFTMAX = 184, REGS_EPILOGUE = 64, REGS_OTHER = 80;
};

// v4 layout (attn_fwd_sm100_v4.cu, tuning builds only): Q 2 tiles | K/V ring (4 x [128 keys][D], K and V tiles interleaved) | one O staging
// tile | row statistics | mbarriers.  TMEM: one shared S buffer, P0, P1, O0, O1.
template <int D>
struct FwdCfg4 {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128 on the tensor-core path");
    static constexpr int NS = (D == 128) ? 4 : 8;               // K/V ring stages
    static constexpr int CHUNKS = D / 64;                       // 128-byte swizzle chunks per row
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;          // [128 rows][128 B]
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_Q = 0;
    static constexpr uint32_t OFF_KV = OFF_Q + 2 * TILE_BYTES;
    static constexpr uint32_t OFF_O = OFF_KV + NS * TILE_BYTES;
    static constexpr uint32_t OFF_STAT = OFF_O + 1 * TILE_BYTES;   // float l[2][128], m[2][128]
    static constexpr uint32_t OFF_WORK = OFF_STAT + 4 * 128 * 4;   // uint32 work_ring[4]
    static constexpr uint32_t OFF_BAR = OFF_WORK + 16;
    static constexpr int NBAR = 29 + 2 * NS;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + NBAR * 8;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
    // TMEM columns (512 allocated)
    static constexpr uint32_t COL_S = 0, COL_P0 = 128, COL_P1 = 192, COL_O0 = 256, COL_O1 = 256 + D;
};

// ---- tcgen05 backward (attn_bwd_sm100.cu) ---------------------------------------------
struct BwdParams {
    void* dq_out;             // dQ kernel: [B,Hq,Sq,D] output in the input dtype
    const void* q;            // dQ kernel: Q and dO [B,Hq,Sq,D] are read straight from global memory into TMEM
    const void* d_o;
    const float* lse;         // [B,Hq,Sq] natural-log LSE of the forward
    const float* delta;       // [B,Hq,Sq] rowsum(O o dO)
    uint32_t B, Hq, Hkv, Sq, Sk;
    uint32_t units_per_run;   // dK/dV kernel: (batch, kv-head) units whose KV-block CTAs are launched together, so that the
                              // Q and dO tiles they all stream stay L2-resident (0 = all units at once)
    float scale, scale_log2;
    int32_t causal;
    int32_t order;            // reserved for A/B tuning of the MMA issue order (unused by the shipped kernels)
    unsigned long long* trace; // bring-up: CTA 0 records (tag << 48 | clock64) events here (3 x 4096 entries) or nullptr
    uint32_t win_left, win_right;   // visibility band i - win_left <= j <= i + win_right (kWinInf = unbounded; causal: win_right = 0)
    uint32_t D_real;          // head_dim of the tensors in memory (<= the kernel's D: TMA zero-fills / clips the padding columns)
    float* dq_acc;            // fused kernel: [B,Hq,Sq,D] fp32 dQ accumulator (zeroed by the host, converted by aule_bwd_dq_convert_*)
};
template <int D>
struct BwdCfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128 on the tensor-core path");
    static constexpr int THREADS = 544;                         // 16 compute warps (4 column quarters x 128 rows) + 1 issuer warp
    static constexpr int CHUNKS = D / 64;
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;          // [128 rows][128 B]
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_K = 0, OFF_V = TILE_BYTES;
    static constexpr uint32_t OFF_Q = 2 * TILE_BYTES;           // Q double-buffered
    static constexpr uint32_t OFF_DO = 4 * TILE_BYTES;          // dO double-buffered
    static constexpr uint32_t OFF_P = 6 * TILE_BYTES;           // P  [128 q rows][128 keys] 16-bit
    // D = 128: 6 operand tiles are 192 KB, so P(i) and dS(i) take turns in ONE 32 KB buffer (dS(i) is written after
    // dV(i) has read P(i); P(i+1) after dK(i) has read dS(i)).  D = 64 has room for both.
    static constexpr bool SHARE_PDS = (D == 128);
    static constexpr uint32_t OFF_DS = SHARE_PDS ? OFF_P : OFF_P + 2 * CHUNK_BYTES;   // dS same shape
    static constexpr uint32_t OFF_BAR = OFF_DS + 2 * CHUNK_BYTES;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + 128;          // 15 mbarriers
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
};

// dK/dV kernel, transposed form (S^T = K Q^T, dP^T = V dO^T; P^T and dS^T stay in TMEM as A operands):
// K_j | V_j | Q ring (3) | dO ring (2) | row statistics (2 x [lse2 128 | delta 128]) | barriers.  No P / dS staging.
template <int D>
struct BwdTCfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128 on the tensor-core path");
    static constexpr int THREADS = 544;
    static constexpr int NQ = 3, NDO = 2;
    static constexpr int CHUNKS = D / 64;
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_K = 0, OFF_V = TILE_BYTES;
    static constexpr uint32_t OFF_Q = 2 * TILE_BYTES;
    static constexpr uint32_t OFF_DO = (2 + NQ) * TILE_BYTES;
    static constexpr uint32_t OFF_STAT = (2 + NQ + NDO) * TILE_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_STAT + 2 * 256 * 4;   // two statistics buffers (published one step ahead)
    static constexpr uint32_t BAR_BYTES = 176;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + BAR_BYTES;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};

// Fused backward (one kernel: dK, dV accumulate in TMEM, dQ partials are reduced into an fp32 accumulator in global memory):
// K_j | V_j | Q ring (2) | dO ring (2) | dS^T tile | column statistics (2 x [lse2 128 | delta 128]) | barriers.
template <int D>
struct BwdFCfg {
    static_assert(D == 128, "the fused backward is written for head_dim 128");
    static constexpr int THREADS = 544;
    static constexpr int NQ = 2, NDO = 2;
    static constexpr int CHUNKS = D / 64;
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_K = 0, OFF_V = TILE_BYTES;
    static constexpr uint32_t OFF_Q = 2 * TILE_BYTES;
    static constexpr uint32_t OFF_DO = (2 + NQ) * TILE_BYTES;
    static constexpr uint32_t OFF_DS = (2 + NQ + NDO) * TILE_BYTES;       // dS^T [128 keys][128 queries] 16-bit, two swizzled chunks
    static constexpr uint32_t OFF_STAT = OFF_DS + 2 * CHUNK_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_STAT + 2 * 256 * 4;
    static constexpr uint32_t BAR_BYTES = 192;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + BAR_BYTES;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};

// (A/B at build time: -DAULE_F2_NQ=4 -DAULE_F2_NDO=3 -DAULE_F2_STG=1 trades the second staging tile for deeper Q / dO rings)
#ifndef AULE_F2_NQ
#define AULE_F2_NQ 3
#define AULE_F2_NDO 2
#define AULE_F2_STG 2
#endif
// Fused backward in 64-query half steps (attn_bwd_fused2_sm100.cu): the dQ partial of a half step is a [64 q][128 d] fp32
// tile that is staged in shared memory and added into the global accumulator by ONE asynchronous TMA bulk reduction.
// K_j | V_j | Q ring (3 x [64][128]) | dO ring (2) | dS^T tile [128 keys][64 q] | 2 staging tiles (32 KB each) | statistics | barriers.
template <int D>
struct BwdF2Cfg {
    static_assert(D == 128, "the fused backward is written for head_dim 128");
    static constexpr int THREADS = 736;                         // 16 P / dS warps + 4 drain warps + issuer, reducer, fence-helper warps
    static constexpr int NQ = AULE_F2_NQ, NDO = AULE_F2_NDO;
    static constexpr int STG_BUFS = AULE_F2_STG;                // fp32 staging tiles (32 KB each)
    static constexpr int QROWS = 64;                            // queries per half step
    static constexpr uint32_t KV_CHUNK_BYTES = 128 * 128;       // [128 keys][128 B]
    static constexpr uint32_t KV_TILE_BYTES = 2 * KV_CHUNK_BYTES;
    static constexpr uint32_t Q_CHUNK_BYTES = QROWS * 128;      // [64 queries][128 B]
    static constexpr uint32_t Q_TILE_BYTES = 2 * Q_CHUNK_BYTES;
    static constexpr uint32_t DS_BYTES = 128 * 128;             // [128 keys][64 q] 16-bit: one swizzled chunk
    static constexpr uint32_t STG_BYTES = QROWS * D * 4;        // [64 q][128 d] fp32, plain row-major (512 B per query row)
    static constexpr uint32_t OFF_K = 0, OFF_V = KV_TILE_BYTES;
    static constexpr uint32_t OFF_Q = 2 * KV_TILE_BYTES;
    static constexpr uint32_t OFF_DO = OFF_Q + NQ * Q_TILE_BYTES;
    static constexpr uint32_t OFF_DS = OFF_DO + NDO * Q_TILE_BYTES;
    static constexpr uint32_t OFF_STG = OFF_DS + DS_BYTES;
    static constexpr uint32_t OFF_STAT = OFF_STG + STG_BUFS * STG_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_STAT + 2 * 256 * 4;
    static constexpr uint32_t BAR_BYTES = 256;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + BAR_BYTES;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};

// dQ kernel (query block outer): K ring | V ring | barriers   (Q, dO, dS live in TMEM)
template <int D>
struct BwdDqCfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128 on the tensor-core path");
    static constexpr int THREADS = 544;
    static constexpr int NK = 4, NV = 3;                        // ring stages
    static constexpr int CHUNKS = D / 64;
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_K = 0;
    static constexpr uint32_t OFF_V = NK * TILE_BYTES;
    static constexpr uint32_t OFF_BAR = (NK + NV) * TILE_BYTES;
    static constexpr uint32_t BAR_BYTES = 192;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + BAR_BYTES;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
};

// dQ kernel v2: K ring (4) | V ring (2) | dO_i tile | barriers   (Q and dS live in TMEM, dO is a shared-memory A operand)
template <int D>
struct BwdDq2Cfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128 on the tensor-core path");
    static constexpr int THREADS = 608;                         // 16 compute warps + 3 issuer warps (S, dP, dQ streams)
    static constexpr int NK = 4, NV = 2;
    static constexpr int CHUNKS = D / 64;
    static constexpr uint32_t CHUNK_BYTES = 128 * 128;
    static constexpr uint32_t TILE_BYTES = CHUNKS * CHUNK_BYTES;
    static constexpr uint32_t OFF_K = 0;
    static constexpr uint32_t OFF_V = NK * TILE_BYTES;
    static constexpr uint32_t OFF_DO = (NK + NV) * TILE_BYTES;
    static constexpr uint32_t OFF_BAR = (NK + NV + 1) * TILE_BYTES;
    static constexpr uint32_t BAR_BYTES = 192;
    static constexpr uint32_t OFF_TMEM_SLOT = OFF_BAR + BAR_BYTES;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM_SLOT + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};


// ---- paged-KV decode (attn_paged_sm100.cu) --------------------------------------------
// One query token per sequence against a block-table KV cache (vLLM layout):
//   q [B,Hq,D] | k_cache, v_cache [num_blocks, block_size, Hkv, D] | block_tables [B, max_blocks] i32 | context_lens [B] i32
struct PagedParams {
    const void* q;               // [B,Hq,D]
    void* out;                   // [B,Hq,D]
    const int32_t* block_tables; // [B, max_blocks]
    const int32_t* context_lens; // [B]
    float* ws_o;                 // [B,Hq,nsplit,D] fp32 unnormalised partial outputs (nsplit > 1)
    float* ws_ml;                // [B,Hq,nsplit,2] fp32 (running max in the log2 domain, partial row sum)
    uint32_t B, Hq, Hkv;
    uint32_t block_size;         // tokens per page, a multiple of 16
    uint32_t max_blocks;         // row pitch of block_tables
    uint32_t nsplit;             // CTAs sharing one (sequence, kv head)
    float scale_log2;            // softmax scale * log2(e)
    int32_t window;              // > 0: only the last `window` tokens of the context are visible; <= 0: all
};
template <int D>
struct PagedCfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128 on the paged decode path");
    static constexpr int TOK = 16;                              // tokens per pipeline stage (one MMA k-step of P.V)
    static constexpr int NS = (D == 128) ? 8 : 16;              // ring stages (K tile + V tile each)
    static constexpr int CONSUMERS = 4;                         // MMA/softmax warps; warp 4 is the TMA producer
    // Tile i of a CTA lives in stage i % NS and belongs to warp i % CONSUMERS.  NS must be a multiple of CONSUMERS so
    // that successive fills of a stage belong to the SAME warp: a warp then reaches "fill k of stage s" only after it
    // consumed fill k-1, and its parity wait cannot be fooled by a fill that is still in flight (TMA tiles land out of
    // order; with NS = 10 a warp saw "phase parity differs" on a stage whose previous fill had not landed yet and read
    // stale data -- reproduced in tools/paged_ring_sim.py, then on the GPU).
    static_assert(NS % CONSUMERS == 0, "ring stages must be a multiple of the consumer warps");
    static constexpr int THREADS = (CONSUMERS + 1) * 32;
    static constexpr uint32_t TILE_BYTES = TOK * D * 2;         // one K or V tile: [16 tokens][D/64 halves][128 B], 128B swizzle
    static constexpr uint32_t STAGE_BYTES = 2 * TILE_BYTES;
    static constexpr uint32_t OFF_RING = 0;
    static constexpr uint32_t OFF_RED_O = NS * STAGE_BYTES;     // float [CONSUMERS][16 heads][D]
    static constexpr uint32_t OFF_RED_ML = OFF_RED_O + CONSUMERS * 16 * D * 4;   // float [CONSUMERS][16][2]
    static constexpr uint32_t OFF_BAR = OFF_RED_ML + CONSUMERS * 16 * 2 * 4;     // full[NS], empty[NS]
    static constexpr uint32_t SMEM_BYTES = OFF_BAR + 2 * NS * 8 + 1024;          // + slack to align the ring to 1024 B
};

}  // namespace aule_kp
