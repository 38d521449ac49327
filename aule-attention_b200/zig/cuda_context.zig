//! CUDA device runtime for the B200 engine -- Zig mirror of the reference's device runtime
//! (src/vulkan_context.zig:52-251 VulkanContext, src/buffer_manager.zig:8-142, src/gpu_tensor.zig:9-108),
//! written against the extern-C shim exported by libaule.so (include/aule.h).  The shim drives the
//! CUDA Driver API (cuInit / primary context / cuModuleLoadData of the embedded sm_100a cubin,
//! csrc/host/engine.cpp), the way src/backends/hip.zig:126-240 drives hipModule*.
//!
//! NOTE: there is no Zig toolchain in the build image; these sources are authored against the
//! same C ABI the tested C++/Python hosts use and are compiled only by zig/build.sh where `zig`
//! is available.
const std = @import("std");

pub const c = struct {
    pub extern "c" fn aule_init() i32;
    pub extern "c" fn aule_shutdown() void;
    pub extern "c" fn aule_get_error() [*:0]const u8;
    pub extern "c" fn aule_get_device_name(buffer: [*]u8, buffer_len: u32) i32;
    pub extern "c" fn aule_device_count() i32;
    pub extern "c" fn aule_get_sm_count(device: i32) i32;
    pub extern "c" fn aule_synchronize(device: i32) i32;
    pub extern "c" fn aule_supports_backward() i32;
    pub extern "c" fn aule_tensor_create(b: u32, h: u32, s: u32, d: u32) u64;
    pub extern "c" fn aule_tensor_destroy(handle: u64) void;
    pub extern "c" fn aule_tensor_upload(handle: u64, data: [*]const f32, count: u32) i32;
    pub extern "c" fn aule_tensor_download(handle: u64, out: [*]f32, count: u32) i32;
    pub extern "c" fn aule_tensor_size(handle: u64) u32;
    pub extern "c" fn aule_attention_forward_gpu(q: u64, k: u64, v: u64, o: u64, rot_cos: u64, rot_sin: u64, causal: i32, window: i32) i32;
    pub extern "c" fn aule_attention_forward_with_lse(q: [*]const f32, k: [*]const f32, v: [*]const f32, o: [*]f32, lse: [*]f32, b: u32, h: u32, s: u32, d: u32, causal: i32) i32;
    pub extern "c" fn aule_attention_backward(q: [*]const f32, k: [*]const f32, v: [*]const f32, o: [*]const f32, d_o: [*]const f32, lse: [*]const f32, dq: [*]f32, dk: [*]f32, dv: [*]f32, b: u32, h: u32, s: u32, d: u32, causal: i32) i32;
    pub extern "c" fn aule_attention_forward_dptr(q: u64, k: u64, v: u64, o: u64, lse_or_0: u64, B: u32, Hq: u32, Hkv: u32, Sq: u32, Sk: u32, D: u32, dtype: i32, scale: f32, causal: i32, window: i32, device: i32, cu_stream: u64) i32;
    pub extern "c" fn aule_attention_backward_dptr(q: u64, k: u64, v: u64, o: u64, d_o: u64, lse: u64, dq: u64, dk: u64, dv: u64, B: u32, Hq: u32, Hkv: u32, Sq: u32, Sk: u32, D: u32, dtype: i32, scale: f32, causal: i32, device: i32, cu_stream: u64) i32;
    pub extern "c" fn aule_attention_forward_paged(q: u64, k: u64, v: u64, o: u64, rot_cos: u64, rot_sin: u64, causal: i32, window: i32) i32;
    pub extern "c" fn aule_attention_paged_decode_dptr(q: u64, k_cache: u64, v_cache: u64, block_tables: u64, context_lens: u64, out: u64, B: u32, Hq: u32, Hkv: u32, D: u32, num_blocks: u32, block_size: u32, max_blocks_per_seq: u32, max_context_len: u32, dtype: i32, scale: f32, window: i32, device: i32, cu_stream: u64) i32;
    pub extern "c" fn aule_rope_dptr(x: u64, out: u64, cos: u64, sin: u64, B: u32, H: u32, S: u32, D: u32, table_rows: u32, interleaved: i32, dtype: i32, inverse: i32, device: i32, cu_stream: u64) i32;
    pub extern "c" fn aule_attention_forward_rope_dptr(q: u64, k: u64, v: u64, o: u64, lse_or_0: u64, cos: u64, sin: u64, table_rows: u32, interleaved: i32, B: u32, Hq: u32, Hkv: u32, Sq: u32, Sk: u32, D: u32, dtype: i32, scale: f32, causal: i32, window: i32, device: i32, cu_stream: u64) i32;
    pub extern "c" fn aule_attention_forward_spanning_dptr(q: u64, k: u64, v: u64, o: u64, lse_or_0: u64, B: u32, Hq: u32, Hkv: u32, Sq: u32, Sk: u32, D: u32, dtype: i32, scale: f32, causal: i32, window: i32, src_device: i32, cu_stream: u64, devices: [*]const i32, num_devices: i32, chunks: i32, timings_ms_or_null: ?[*]f32) i32;
    pub extern "c" fn aule_smoke_multiply(in: [*]const f32, out: [*]f32, n: u32) i32;
};

pub const CudaError = error{ InitFailed, OutOfDeviceMemory, TransferFailed, ComputeFailed, InvalidShape };

pub const DType = enum(i32) { f32 = 0, bf16 = 1, f16 = 2 };

/// Mirror of VulkanContext (vulkan_context.zig:52-251): owns the device runtime.
pub const CudaContext = struct {
    device_count: u32,
    sm_count: u32,
    device_name: [256]u8,

    pub fn init() CudaError!CudaContext {
        if (c.aule_init() != 0) {
            std.log.err("aule: {s}", .{c.aule_get_error()});
            return CudaError.InitFailed;
        }
        var self = CudaContext{ .device_count = @intCast(c.aule_device_count()), .sm_count = @intCast(c.aule_get_sm_count(0)), .device_name = undefined };
        _ = c.aule_get_device_name(&self.device_name, self.device_name.len);
        return self;
    }

    pub fn deinit(self: *CudaContext) void {
        _ = self;
        c.aule_shutdown();
    }

    /// vulkan_context.zig:151 waitIdle
    pub fn waitIdle(self: *const CudaContext) void {
        _ = self;
        _ = c.aule_synchronize(0);
    }
};

/// Mirror of GpuTensor (gpu_tensor.zig:9-108): shape + device buffer + upload/download.
/// Storage is true HBM here (the reference allocates host-visible memory, gpu_tensor.zig:50).
pub const GpuTensor = struct {
    handle: u64,
    shape: [4]u32,
    element_count: u32,

    pub fn init(shape: [4]u32) CudaError!GpuTensor {
        const h = c.aule_tensor_create(shape[0], shape[1], shape[2], shape[3]);
        if (h == 0) return CudaError.OutOfDeviceMemory;
        return GpuTensor{ .handle = h, .shape = shape, .element_count = shape[0] * shape[1] * shape[2] * shape[3] };
    }
    pub fn deinit(self: *GpuTensor) void {
        c.aule_tensor_destroy(self.handle);
        self.handle = 0;
    }
    pub fn upload(self: *const GpuTensor, data: []const f32) CudaError!void {
        if (data.len != self.element_count) return CudaError.InvalidShape; // gpu_tensor.zig:71-73
        if (c.aule_tensor_upload(self.handle, data.ptr, @intCast(data.len)) != 0) return CudaError.TransferFailed;
    }
    pub fn download(self: *const GpuTensor, out: []f32) CudaError!void {
        if (out.len != self.element_count) return CudaError.InvalidShape; // gpu_tensor.zig:84-86
        if (c.aule_tensor_download(self.handle, out.ptr, @intCast(out.len)) != 0) return CudaError.TransferFailed;
    }
};
