//! Mirror of src/attention_pipeline.zig:10-392 (AttentionPushConstants, AttentionPipeline) and of
//! src/attention_backward_pipeline.zig:8-538 (ForwardWithLsePipeline, BackwardPipeline).
//! A "pipeline" here is a launch description for the sm_100a kernels: the push-constant block
//! becomes the argument list of aule_attention_{forward,backward}_dptr; descriptor sets become raw
//! device pointers; vkCmdDispatch + fence (attention_pipeline.zig:372-390) becomes one asynchronous
//! launch on a CUDA stream (no per-call blocking fence).
const std = @import("std");
const cuda = @import("cuda_context.zig");

/// attention_pipeline.zig:10-21
pub const AttentionPushConstants = extern struct {
    batch_size: u32,
    num_heads: u32, // q heads
    seq_len: u32, // q sequence length
    head_dim: u32,
    scale: f32, // <= 0 => 1/sqrt(head_dim), computed in the engine like attention_pipeline.zig:329
    causal: u32,
    has_rope: u32 = 0, // RoPE is a prologue launch of the C side (aule_attention_forward_rope_dptr), not a push constant
    num_kv_heads: u32,
    key_seq_len: u32,
    window_size: i32 = -1,
};

pub const DeviceBuffers = struct { q: u64, k: u64, v: u64, o: u64, lse: u64 = 0 };

pub const AttentionPipeline = struct {
    dtype: cuda.DType,
    device: i32 = 0,
    stream: u64 = 0,
    bufs: DeviceBuffers = .{ .q = 0, .k = 0, .v = 0, .o = 0 },

    pub fn init(dtype: cuda.DType) AttentionPipeline {
        return .{ .dtype = dtype };
    }
    pub fn deinit(self: *AttentionPipeline) void {
        _ = self;
    }
    /// attention_pipeline.zig:217 updateDescriptors
    pub fn updateDescriptors(self: *AttentionPipeline, bufs: DeviceBuffers) void {
        self.bufs = bufs;
    }
    /// attention_pipeline.zig:312 dispatch. Grid selection lives in the engine: a persistent
    /// grid of min(work items, SM count) CTAs instead of (1, ceil(S/16), B*H) workgroups (:341-372).
    pub fn dispatch(self: *AttentionPipeline, pc: AttentionPushConstants) cuda.CudaError!void {
        const rc = cuda.c.aule_attention_forward_dptr(self.bufs.q, self.bufs.k, self.bufs.v, self.bufs.o, self.bufs.lse, pc.batch_size, pc.num_heads, pc.num_kv_heads, pc.seq_len, pc.key_seq_len, pc.head_dim, @intFromEnum(self.dtype), pc.scale, @intCast(pc.causal), pc.window_size, self.device, self.stream);
        if (rc != 0) {
            std.log.err("attention dispatch: {s}", .{cuda.c.aule_get_error()});
            return cuda.CudaError.ComputeFailed;
        }
    }
};

/// attention_backward_pipeline.zig:29-276 -- the forward that also stores LSE is the same kernel
/// with a non-null `lse` pointer.
pub const ForwardWithLsePipeline = AttentionPipeline;

pub const BackwardBuffers = struct { q: u64, k: u64, v: u64, o: u64, d_o: u64, lse: u64, dq: u64, dk: u64, dv: u64 };

/// attention_backward_pipeline.zig:279-538
pub const BackwardPipeline = struct {
    dtype: cuda.DType,
    device: i32 = 0,
    stream: u64 = 0,
    bufs: ?BackwardBuffers = null,

    pub fn init(dtype: cuda.DType) BackwardPipeline {
        return .{ .dtype = dtype };
    }
    pub fn updateDescriptors(self: *BackwardPipeline, bufs: BackwardBuffers) void {
        self.bufs = bufs;
    }
    pub fn dispatch(self: *BackwardPipeline, pc: AttentionPushConstants) cuda.CudaError!void {
        const b = self.bufs orelse return cuda.CudaError.InvalidShape;
        const rc = cuda.c.aule_attention_backward_dptr(b.q, b.k, b.v, b.o, b.d_o, b.lse, b.dq, b.dk, b.dv, pc.batch_size, pc.num_heads, pc.num_kv_heads, pc.seq_len, pc.key_seq_len, pc.head_dim, @intFromEnum(self.dtype), pc.scale, @intCast(pc.causal), self.device, self.stream);
        if (rc != 0) return cuda.CudaError.ComputeFailed;
    }
};
