//! Mirror of src/attention_gpu.zig:31-830 (AttentionEngine) on the CUDA engine: same method names and
//! validation rules (attention_gpu.zig:372-404, widened to D <= 128, GQA, Sq != Sk).
const std = @import("std");
const cuda = @import("cuda_context.zig");
const pipe = @import("attention_pipeline.zig");

pub const AttentionEngine = struct {
    ctx: cuda.CudaContext,
    backward_supported: bool,

    /// attention_gpu.zig:63 initWithBackward
    pub fn initWithBackward() cuda.CudaError!AttentionEngine {
        const ctx = try cuda.CudaContext.init();
        return .{ .ctx = ctx, .backward_supported = cuda.c.aule_supports_backward() == 1 };
    }
    /// attention_gpu.zig:324 deinit
    pub fn deinit(self: *AttentionEngine) void {
        self.ctx.deinit();
    }
    /// attention_gpu.zig:347
    pub fn supportsBackward(self: *const AttentionEngine) bool {
        return self.backward_supported;
    }
    /// attention_gpu.zig:352 createTensor
    pub fn createTensor(self: *AttentionEngine, shape: [4]u32) cuda.CudaError!cuda.GpuTensor {
        _ = self;
        return cuda.GpuTensor.init(shape);
    }
    /// attention_gpu.zig:472 synchronize
    pub fn synchronize(self: *AttentionEngine) void {
        self.ctx.waitIdle();
    }

    fn validate(q: *const cuda.GpuTensor, k: *const cuda.GpuTensor, v: *const cuda.GpuTensor, o: *const cuda.GpuTensor) cuda.CudaError!void {
        if (k.shape[0] != q.shape[0] or k.shape[3] != q.shape[3]) return cuda.CudaError.InvalidShape; // :376-381
        if (q.shape[1] % k.shape[1] != 0) return cuda.CudaError.InvalidShape; // :383-388 GQA divisibility
        inline for (0..4) |i| {
            if (v.shape[i] != k.shape[i]) return cuda.CudaError.InvalidShape; // :389-394
            if (o.shape[i] != q.shape[i]) return cuda.CudaError.InvalidShape; // :396-400
        }
        if (q.shape[3] > 128 or q.shape[3] % 4 != 0) return cuda.CudaError.InvalidShape; // :402-404 (64 there)
    }

    /// attention_gpu.zig:360 forward (fp32 handle tensors; asynchronous in the reference's sense: the
    /// C entry synchronises like forwardSync, attention_gpu.zig:456-469)
    pub fn forward(self: *AttentionEngine, q: *const cuda.GpuTensor, k: *const cuda.GpuTensor, v: *const cuda.GpuTensor, o: *const cuda.GpuTensor, causal: bool, window_size: i32) cuda.CudaError!void {
        _ = self;
        try validate(q, k, v, o);
        if (cuda.c.aule_attention_forward_gpu(q.handle, k.handle, v.handle, o.handle, 0, 0, @intFromBool(causal), window_size) != 0)
            return cuda.CudaError.ComputeFailed;
    }
    pub fn forwardSync(self: *AttentionEngine, q: *const cuda.GpuTensor, k: *const cuda.GpuTensor, v: *const cuda.GpuTensor, o: *const cuda.GpuTensor, causal: bool) cuda.CudaError!void {
        try self.forward(q, k, v, o, causal, -1);
        self.synchronize();
    }

    /// Raw device pointers of any supported dtype on a caller stream -- the path PyTorch tensors take.
    pub fn forwardDevice(self: *AttentionEngine, dtype: cuda.DType, bufs: pipe.DeviceBuffers, pc: pipe.AttentionPushConstants, stream: u64) cuda.CudaError!void {
        _ = self;
        var p = pipe.AttentionPipeline.init(dtype);
        p.stream = stream;
        p.updateDescriptors(bufs);
        try p.dispatch(pc);
    }

    /// attention_gpu.zig:484-653 forwardPaged: same tensors and result as forward(); the reference pages K/V into a
    /// private block pool first, the B200 library runs the fused kernel on them directly.
    pub fn forwardPaged(self: *AttentionEngine, q: *const cuda.GpuTensor, k: *const cuda.GpuTensor, v: *const cuda.GpuTensor, o: *const cuda.GpuTensor, causal: bool, window_size: i32) cuda.CudaError!void {
        _ = self;
        if (cuda.c.aule_attention_forward_paged(q.handle, k.handle, v.handle, o.handle, 0, 0, @intFromBool(causal), window_size) != 0)
            return cuda.CudaError.ComputeFailed;
    }

    /// Serving-side paged KV cache (caller-owned block tables, vLLM layout): one query token per sequence.
    /// q/out [B,Hq,D]; k_cache/v_cache [num_blocks, block_size, Hkv, D]; block_tables [B, max_blocks] i32; context_lens [B] i32.
    pub fn pagedDecodeDevice(self: *AttentionEngine, dtype: cuda.DType, q: u64, k_cache: u64, v_cache: u64, block_tables: u64, context_lens: u64, out: u64, dims: struct { B: u32, Hq: u32, Hkv: u32, D: u32, num_blocks: u32, block_size: u32, max_blocks: u32, max_context: u32 }, scale: f32, window: i32, device: i32, stream: u64) cuda.CudaError!void {
        _ = self;
        if (cuda.c.aule_attention_paged_decode_dptr(q, k_cache, v_cache, block_tables, context_lens, out, dims.B, dims.Hq, dims.Hkv, dims.D, dims.num_blocks, dims.block_size, dims.max_blocks, dims.max_context, @intFromEnum(dtype), scale, window, device, stream) != 0)
            return cuda.CudaError.ComputeFailed;
    }

    /// attention_gpu.zig:707 forwardWithLse (host fp32 slices, MHA)
    pub fn forwardWithLse(self: *AttentionEngine, q: []const f32, k: []const f32, v: []const f32, o: []f32, lse: []f32, shape: [4]u32, causal: bool) cuda.CudaError!void {
        _ = self;
        if (cuda.c.aule_attention_forward_with_lse(q.ptr, k.ptr, v.ptr, o.ptr, lse.ptr, shape[0], shape[1], shape[2], shape[3], @intFromBool(causal)) != 0)
            return cuda.CudaError.ComputeFailed;
    }
    /// attention_gpu.zig:771 backward / :815 backwardSync
    pub fn backwardSync(self: *AttentionEngine, q: []const f32, k: []const f32, v: []const f32, o: []const f32, d_o: []const f32, lse: []const f32, dq: []f32, dk: []f32, dv: []f32, shape: [4]u32, causal: bool) cuda.CudaError!void {
        if (!self.backward_supported) return cuda.CudaError.ComputeFailed;
        if (cuda.c.aule_attention_backward(q.ptr, k.ptr, v.ptr, o.ptr, d_o.ptr, lse.ptr, dq.ptr, dk.ptr, dv.ptr, shape[0], shape[1], shape[2], shape[3], @intFromBool(causal)) != 0)
            return cuda.CudaError.ComputeFailed;
    }
};

test "known answer: Q=K=0.5, V=[[1,2,3,4],[5,6,7,8]] -> [3,4,5,6] (src/attention_ref.zig:250-298)" {
    var eng = AttentionEngine.initWithBackward() catch return; // no B200: skip
    defer eng.deinit();
    const shape = [4]u32{ 1, 1, 2, 4 };
    var q = try eng.createTensor(shape);
    defer q.deinit();
    var k = try eng.createTensor(shape);
    defer k.deinit();
    var v = try eng.createTensor(shape);
    defer v.deinit();
    var o = try eng.createTensor(shape);
    defer o.deinit();
    const half = [_]f32{0.5} ** 8;
    const vals = [_]f32{ 1, 2, 3, 4, 5, 6, 7, 8 };
    try q.upload(&half);
    try k.upload(&half);
    try v.upload(&vals);
    try eng.forwardSync(&q, &k, &v, &o, false);
    var out: [8]f32 = undefined;
    try o.download(&out);
    const expect = [_]f32{ 3, 4, 5, 6, 3, 4, 5, 6 };
    for (out, expect) |a, e| try std.testing.expect(@abs(a - e) < 1e-3);
}
