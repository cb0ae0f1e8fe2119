//! Device-resident, P-sharded path: ri3ao stays in HBM across SCF iterations; one `Context` per GPU (one process or one
//! thread per GPU).  The partition is the reference's `iter_auxbas(P_lo..P_hi)` (src/ri.rs:190-198); the only exchanges
//! are one all-reduce each for J and K (NCCL inside librest_b200, bound at run time).
use crate::ffi::*;
use std::ffi::{c_int, c_void};
use std::ptr;

pub struct Context { pub raw: *mut RbCtx, pub device: i32 }

impl Context {
    pub fn new(device: i32) -> Context {
        let mut raw: *mut RbCtx = ptr::null_mut();
        unsafe { check(rb_ctx_create(device as c_int, &mut raw), "rb_ctx_create"); }
        Context { raw, device }
    }
    pub fn sync(&self) { unsafe { check(rb_ctx_sync(self.raw), "rb_ctx_sync"); } }
    /// rank 0 creates the id and hands the 128 bytes to the other ranks by the host's own means
    pub fn unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        unsafe { check(rb_comm_unique_id(id.as_mut_ptr()), "rb_comm_unique_id"); }
        id
    }
    pub fn comm_init(&self, rank: usize, world: usize, id: &[u8; 128]) {
        unsafe { check(rb_comm_init_rank(self.raw, rank as c_int, world as c_int, id.as_ptr()), "rb_comm_init_rank"); }
    }
}

impl Drop for Context {
    fn drop(&mut self) { unsafe { rb_ctx_destroy(self.raw); } }
}

/// A recorded call sequence (CUDA graph): the d_P / J / K calls of one SCF iteration issued once between `Context::record`
/// and `Recording::finish`, then replayed with one launch per iteration.  See include/rest_b200.h (rb_graph_begin) for the
/// rules (non-default stream, one un-recorded pass first, no eigen-solver / collective inside).
pub struct Graph<'a> { ctx: &'a Context, raw: *mut c_void }

impl Context {
    /// the library's own stream: the default stream cannot be captured
    pub fn use_own_stream(&self) { unsafe { check(rb_ctx_use_own_stream(self.raw), "rb_ctx_use_own_stream"); } }
    /// record the calls `f` issues on this context; they do not run
    pub fn record<F: FnOnce()>(&self, f: F) -> Graph<'_> {
        let mut raw: *mut c_void = ptr::null_mut();
        unsafe { check(rb_graph_begin(self.raw), "rb_graph_begin"); }
        f();
        unsafe { check(rb_graph_end(self.raw, &mut raw), "rb_graph_end"); }
        Graph { ctx: self, raw }
    }
}

impl<'a> Graph<'a> {
    pub fn launch(&self) { unsafe { check(rb_graph_launch(self.ctx.raw, self.raw), "rb_graph_launch"); } }
    pub fn kernels(&self) -> i64 { unsafe { rb_graph_kernel_count(self.raw) } }
}

impl<'a> Drop for Graph<'a> {
    fn drop(&mut self) { unsafe { rb_graph_free(self.ctx.raw, self.raw); } }
}

/// FP64 buffer in HBM owned by a context
pub struct DeviceBuffer<'a> { pub ctx: &'a Context, pub ptr: *mut f64, pub len: usize }

impl<'a> DeviceBuffer<'a> {
    pub fn new(ctx: &'a Context, len: usize) -> DeviceBuffer<'a> {
        let mut p: *mut c_void = ptr::null_mut();
        unsafe { check(rb_dev_alloc(ctx.raw, (len.max(1) * 8) as i64, &mut p), "rb_dev_alloc"); }
        DeviceBuffer { ctx, ptr: p as *mut f64, len }
    }
    pub fn upload(&self, src: &[f64]) {
        assert!(src.len() <= self.len);
        unsafe { check(rb_memcpy_h2d(self.ctx.raw, self.ptr as *mut c_void, src.as_ptr() as *const c_void, (src.len() * 8) as i64), "rb_memcpy_h2d"); }
    }
    pub fn download(&self, dst: &mut [f64]) {
        assert!(dst.len() <= self.len);
        unsafe { check(rb_memcpy_d2h(self.ctx.raw, dst.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, (dst.len() * 8) as i64), "rb_memcpy_d2h"); }
        self.ctx.sync();
    }
}

impl<'a> Drop for DeviceBuffer<'a> {
    fn drop(&mut self) { unsafe { rb_dev_free(self.ctx.raw, self.ptr as *mut c_void); } }
}

/// the slab range of rank r of `world`: floor(r naux / G) .. floor((r+1) naux / G)
pub fn shard_range(naux: usize, rank: usize, world: usize) -> std::ops::Range<usize> {
    (rank * naux / world)..((rank + 1) * naux / world)
}

/// One rank's slabs of ri3ao[nb, nb, naux] in HBM
pub struct ShardedRI<'a> { pub ctx: &'a Context, pub nb: usize, pub naux: usize, pub p: std::ops::Range<usize>, pub data: DeviceBuffer<'a> }

impl<'a> ShardedRI<'a> {
    /// upload this rank's slabs from the full host tensor (the chunk `iter_auxbas(p)` yields)
    pub fn from_host(ctx: &'a Context, ri3ao: &[f64], nb: usize, naux: usize, rank: usize, world: usize) -> ShardedRI<'a> {
        let p = shard_range(naux, rank, world);
        let data = DeviceBuffer::new(ctx, nb * nb * p.len());
        data.upload(&ri3ao[nb * nb * p.start..nb * nb * p.end]);
        ShardedRI { ctx, nb, naux, p, data }
    }
    fn nx(&self) -> c_int { self.p.len() as c_int }
    /// this rank's rows of ri3mo[P, a, b] (P-fastest, pitch = local slab count); no communication
    pub fn ao2mo(&self, c_left: &DeviceBuffer, nl: usize, c_right: &DeviceBuffer, nr: usize, out: &DeviceBuffer) {
        unsafe {
            check(rb_ri_ao2mo(self.ctx.raw, c_left.ptr, nl as c_int, c_right.ptr, nr as c_int, self.data.ptr, out.ptr, self.nb as c_int,
                              self.nx(), self.p.len() as i64), "rb_ri_ao2mo");
        }
    }
    /// d[P] for the local P (D replicated); `gather` = the full vector on every rank
    pub fn dp(&self, dm: &DeviceBuffer, d_local: &DeviceBuffer) {
        unsafe { check(rb_ri_dp(self.ctx.raw, self.data.ptr, dm.ptr, d_local.ptr, self.nb as c_int, self.nx()), "rb_ri_dp"); }
    }
    pub fn gather_dp(&self, d_local: &DeviceBuffer, d_full: &DeviceBuffer) {
        unsafe { check(rb_allgather_shards(self.ctx.raw, d_local.ptr, d_full.ptr, self.naux as i64), "rb_allgather_shards"); }
    }
    /// J complete on every rank (partial sum + one all-reduce)
    pub fn j(&self, d_local: &DeviceBuffer, j: &DeviceBuffer) {
        unsafe { check(rb_ri_j_allreduce(self.ctx.raw, self.data.ptr, d_local.ptr, j.ptr, self.nb as c_int, self.nx()), "rb_ri_j_allreduce"); }
    }
    /// d[P_lo..P_hi) and J (complete on every rank) from ONE read of the shard where that is faster, else the two passes
    pub fn dp_j(&self, dm: &DeviceBuffer, d_local: &DeviceBuffer, j: &DeviceBuffer) {
        unsafe {
            check(rb_ri_dp_j(self.ctx.raw, self.data.ptr, dm.ptr, d_local.ptr, j.ptr, self.nb as c_int, self.nx()), "rb_ri_dp_j");
            check(rb_allreduce_sum(self.ctx.raw, j.ptr, (self.nb * self.nb) as i64), "rb_allreduce_sum");
        }
    }
    /// K complete on every rank
    pub fn k(&self, ct: &DeviceBuffer, no: usize, k: &DeviceBuffer) {
        unsafe { check(rb_ri_k_allreduce(self.ctx.raw, self.data.ptr, ct.ptr, no as c_int, k.ptr, self.nb as c_int, self.nx()), "rb_ri_k_allreduce"); }
    }
}
