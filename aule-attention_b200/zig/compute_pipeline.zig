//! Mirror of src/compute_pipeline.zig:7-370 (the generic two-buffer pipeline used by
//! tests/test_multiply.zig with shaders/test.comp): here the x2 kernel `aule_smoke_multiply`
//! of the embedded sm_100a module, launched through the same module-load / launch / copy path
//! as the attention kernels (csrc/host/engine.cpp Engine::smoke_multiply).
const std = @import("std");
const cuda = @import("cuda_context.zig");

pub const ComputePipeline = struct {
    ctx: *const cuda.CudaContext,

    pub fn init(ctx: *const cuda.CudaContext) ComputePipeline {
        return .{ .ctx = ctx };
    }
    pub fn deinit(self: *ComputePipeline) void {
        _ = self;
    }
    /// compute_pipeline.zig:254 recordCopyAndDispatch: out[i] = 2 * in[i]
    pub fn recordCopyAndDispatch(self: *ComputePipeline, input: []const f32, output: []f32) cuda.CudaError!void {
        _ = self;
        if (input.len != output.len) return cuda.CudaError.InvalidShape;
        if (cuda.c.aule_smoke_multiply(input.ptr, output.ptr, @intCast(input.len)) != 0) return cuda.CudaError.ComputeFailed;
    }
};

test "multiply by two (tests/test_multiply.zig analogue)" {
    var ctx = cuda.CudaContext.init() catch return; // no B200: skip like the reference skips without Vulkan
    defer ctx.deinit();
    var pipe = ComputePipeline.init(&ctx);
    var in: [256]f32 = undefined;
    var out: [256]f32 = undefined;
    for (&in, 0..) |*x, i| x.* = @floatFromInt(i);
    try pipe.recordCopyAndDispatch(&in, &out);
    for (in, out) |a, b| try std.testing.expectEqual(2.0 * a, b);
}
