//! `B200Matrix` -- the reference-side binding of libb200spgemm (include/b200_spgemm.h).
//!
//! Drop this file into the reference as `src/graph_b200.rs`, add `pub mod graph_b200;` to
//! `src/lib.rs` (module list, src/lib.rs:5-14) and `rust/build.rs` as the crate's `build.rs`.
//! The type exposes the same inherent-method surface as `MagnusMatrix`
//! (src/graph_magnus.rs:16-448) so that `bench_repeated_exponentiation` (:699-788) and
//! `bench_matmul_magnus` (:790-929) gain one more column by adding
//! `let a_b200 = B200Matrix::from_csr_parts(n, &a_csr.row_ptr, &a_csr.col_idx, &vals64);` and
//! `prev_b200.matmul(&a_b200)`.
//!
//! NOT COMPILED in the build image (no Rust toolchain there); the C ABI below is exercised by
//! the ctypes host in `sparse_linear_algebra_tests_b200/_native.py`, which binds the very same
//! symbols.  Values are `u64` with saturating arithmetic (`Sat64`, src/graph_sprs.rs:15-86);
//! a `u32`-valued twin for `CsrMatrix` only differs in `val_bits`.
#![allow(dead_code)]

use std::cell::OnceCell;
use std::collections::BTreeMap;
use std::ffi::{c_char, c_int, c_void, CStr};
use std::sync::OnceLock;

use einsum_dyn::NDIndex;
use rand::Rng;

#[repr(C)]
pub struct B200Ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct B200Csr {
    _private: [u8; 0],
}
#[repr(C)]
pub struct B200CommRaw {
    _private: [u8; 0],
}

/// Mirror of `b200_stats` (include/b200_spgemm.h).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct B200Stats {
    pub rows: u64,
    pub cols: u64,
    pub nnz_a: u64,
    pub nnz_b: u64,
    pub nnz_c: u64,
    pub products: u64,
    pub max_row_products: u64,
    pub max_row_nnz: u64,
    pub bytes_algorithmic: u64,
    pub ms_symbolic: f32,
    pub ms_numeric: f32,
    pub ms_total: f32,
    pub acc_mode: i32,
    pub kernel_launches: i32,
    pub sym_bin_rows: [u32; 16],
    pub pipeline: u32,
    pub reserved: [u32; 15],
}

pub const B200_OK: c_int = 0;
pub const B200_ERR_SHAPE: c_int = 2;

extern "C" {
    fn b200_last_error() -> *const c_char;
    fn b200_ctx_create(device: c_int, cuda_stream: *mut c_void, out: *mut *mut B200Ctx) -> c_int;
    fn b200_ctx_destroy(ctx: *mut B200Ctx) -> c_int;
    fn b200_ctx_synchronize(ctx: *mut B200Ctx) -> c_int;
    fn b200_csr_upload(
        ctx: *mut B200Ctx, rows: u64, cols: u64, row_ptr: *const u64, col_idx: *const u32,
        values: *const c_void, val_bits: c_int, out: *mut *mut B200Csr,
    ) -> c_int;
    fn b200_csr_upload_idx64(
        ctx: *mut B200Ctx, rows: u64, cols: u64, row_ptr: *const u64, col_idx: *const u64,
        values: *const c_void, val_bits: c_int, out: *mut *mut B200Csr,
    ) -> c_int;
    fn b200_csr_free(ctx: *mut B200Ctx, m: *mut B200Csr) -> c_int;
    fn b200_csr_info(m: *const B200Csr, rows: *mut u64, cols: *mut u64, nnz: *mut u64, val_bits: *mut c_int) -> c_int;
    fn b200_csr_download_idx64(
        ctx: *mut B200Ctx, m: *const B200Csr, row_ptr: *mut u64, col_idx: *mut u64, values: *mut c_void,
    ) -> c_int;
    fn b200_spgemm(
        ctx: *mut B200Ctx, a: *const B200Csr, b: *const B200Csr, c: *mut *mut B200Csr, stats: *mut B200Stats,
    ) -> c_int;
    fn b200_csr_add(ctx: *mut B200Ctx, a: *const B200Csr, b: *const B200Csr, c: *mut *mut B200Csr) -> c_int;
    fn b200_csr_same_pattern(ctx: *mut B200Ctx, a: *const B200Csr, b: *const B200Csr, same: *mut c_int) -> c_int;
    fn b200_csr_row_block(ctx: *mut B200Ctx, a: *const B200Csr, row_begin: u64, row_end: u64, out: *mut *mut B200Csr) -> c_int;
    fn b200_lattice(ctx: *mut B200Ctx, dims: *const u64, ndims: c_int, torus: c_int, val_bits: c_int, out: *mut *mut B200Csr) -> c_int;
    fn b200_thin(
        ctx: *mut B200Ctx, a: *const B200Csr, density: f64, seed32: *const u8, skip_draws: u64, out: *mut *mut B200Csr, draws_consumed: *mut u64,
    ) -> c_int;
    fn b200_csr_from_coo(
        ctx: *mut B200Ctx, rows: u64, cols: u64, n: u64, row_idx: *const u32, col_idx: *const u32, values: *const c_void, val_bits: c_int,
        saturating: c_int, out: *mut *mut B200Csr,
    ) -> c_int;
    fn b200_rmat(ctx: *mut B200Ctx, scale: c_int, edge_factor: u64, a: f64, b: f64, c: f64, seed: u64, val_bits: c_int, out: *mut *mut B200Csr) -> c_int;
    fn b200_csr_bandwidth_stats(ctx: *mut B200Ctx, m: *const B200Csr, max_bw: *mut u64, avg_bw: *mut f64) -> c_int;
    fn b200_csr_permute(ctx: *mut B200Ctx, a: *const B200Csr, perm: *const u32, out: *mut *mut B200Csr) -> c_int;
    fn b200_csr_rcm_order(ctx: *mut B200Ctx, a: *const B200Csr, perm_out: *mut u32) -> c_int;
    fn b200_shard_rows_by_products(ctx: *mut B200Ctx, a: *const B200Csr, b: *const B200Csr, nparts: c_int, cuts: *mut u64) -> c_int;
    // multi-GPU: one communicator per (process, GPU); see include/b200_spgemm.h
    fn b200_comm_unique_id(id128: *mut u8) -> c_int;
    fn b200_comm_init_rank(ctx: *mut B200Ctx, nranks: c_int, rank: c_int, id128: *const u8, out: *mut *mut B200CommRaw) -> c_int;
    fn b200_comm_init_all(ctxs: *mut *mut B200Ctx, ngpus: c_int, out: *mut *mut B200CommRaw) -> c_int;
    fn b200_comm_destroy(c: *mut B200CommRaw) -> c_int;
    fn b200_comm_broadcast_csr(c: *mut B200CommRaw, src: *const B200Csr, root: c_int, out: *mut *mut B200Csr) -> c_int;
    fn b200_comm_allgather_csr(c: *mut B200CommRaw, block: *const B200Csr, out: *mut *mut B200Csr) -> c_int;
    fn b200_comm_allreduce(c: *mut B200CommRaw, host_scalars: *mut c_void, n: c_int, op: c_int) -> c_int;
}

struct Ctx(*mut B200Ctx);
// One engine context per process, used from one thread at a time (the benches are single-threaded
// callers; the engine itself parallelises on the GPU where the reference used the rayon pool).
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

fn ctx() -> *mut B200Ctx {
    static CTX: OnceLock<Ctx> = OnceLock::new();
    CTX.get_or_init(|| {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { b200_ctx_create(0, std::ptr::null_mut(), &mut p) };
        check(rc);
        Ctx(p)
    })
    .0
}

/// The reference panics (`assert_eq!`) where the ABI returns a status: keep that behaviour.
fn check(rc: c_int) {
    if rc != B200_OK {
        let msg = unsafe { CStr::from_ptr(b200_last_error()) }.to_string_lossy().into_owned();
        panic!("b200 engine error {rc}: {msg}");
    }
}

/// n x n saturating-u64 path-count matrix resident on the GPU.
pub struct B200Matrix {
    pub n: usize,
    handle: *mut B200Csr,
    /// Host copy (row_ptr, col_idx, values), fetched on the first `get`/`print`.
    host: OnceCell<(Vec<usize>, Vec<usize>, Vec<u64>)>,
}

impl Drop for B200Matrix {
    fn drop(&mut self) {
        unsafe { b200_csr_free(ctx(), self.handle) };
    }
}

impl Clone for B200Matrix {
    /// Device-to-device copy (the full row range as a "row block").
    fn clone(&self) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_csr_row_block(ctx(), self.handle, 0, self.n as u64, &mut h) });
        Self::wrap(self.n, h)
    }
}

impl B200Matrix {
    fn wrap(n: usize, handle: *mut B200Csr) -> Self {
        Self { n, handle, host: OnceCell::new() }
    }

    /// Upload MAGNUS-layout CSR arrays (`usize` columns, src/graph_magnus.rs:52-74).
    pub fn from_csr_parts_usize(n: usize, row_ptr: &[usize], col_idx: &[usize], values: &[u64]) -> Self {
        assert_eq!(row_ptr.len(), n + 1);
        assert_eq!(col_idx.len(), values.len());
        let rp: Vec<u64> = row_ptr.iter().map(|&x| x as u64).collect();
        let ci: Vec<u64> = col_idx.iter().map(|&x| x as u64).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe {
            b200_csr_upload_idx64(ctx(), n as u64, n as u64, rp.as_ptr(), ci.as_ptr(), values.as_ptr() as *const c_void, 64, &mut h)
        });
        Self::wrap(n, h)
    }

    /// Upload `CsrMatrix`-layout arrays (`u32` columns, src/graph_csr.rs:42-53) with values widened to u64.
    pub fn from_csr_parts(n: usize, row_ptr: &[usize], col_idx: &[u32], values: &[u64]) -> Self {
        assert_eq!(row_ptr.len(), n + 1);
        let rp: Vec<u64> = row_ptr.iter().map(|&x| x as u64).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe {
            b200_csr_upload(ctx(), n as u64, n as u64, rp.as_ptr(), col_idx.as_ptr(), values.as_ptr() as *const c_void, 64, &mut h)
        });
        Self::wrap(n, h)
    }

    fn host(&self) -> &(Vec<usize>, Vec<usize>, Vec<u64>) {
        self.host.get_or_init(|| {
            let nnz = self.nnz();
            let mut rp = vec![0u64; self.n + 1];
            let mut ci = vec![0u64; nnz];
            let mut vv = vec![0u64; nnz];
            check(unsafe {
                b200_csr_download_idx64(ctx(), self.handle, rp.as_mut_ptr(), ci.as_mut_ptr(), vv.as_mut_ptr() as *mut c_void)
            });
            (rp.into_iter().map(|x| x as usize).collect(), ci.into_iter().map(|x| x as usize).collect(), vv)
        })
    }

    pub fn row_ptr(&self) -> &[usize] { &self.host().0 }
    pub fn col_idx(&self) -> &[usize] { &self.host().1 }
    pub fn values(&self) -> &[u64] { &self.host().2 }

    // ------------------------------------------------------------------ builders (host side, like the reference)
    pub fn new(n: usize) -> Self {
        Self::from_csr_parts_usize(n, &vec![0; n + 1], &[], &[])
    }

    pub fn identity(n: usize) -> Self {
        let rp: Vec<usize> = (0..=n).collect();
        let ci: Vec<usize> = (0..n).collect();
        Self::from_csr_parts_usize(n, &rp, &ci, &vec![1u64; n])
    }

    /// Sort by (row, col), sum duplicates with plain `+=`, drop zeros (src/graph_magnus.rs:34-76).
    fn from_coo(n: usize, triplets: &mut Vec<(usize, usize, u64)>) -> Self {
        triplets.sort_unstable_by_key(|t| (t.0, t.1));
        let mut rp = vec![0usize; n + 1];
        let mut ci: Vec<usize> = Vec::with_capacity(triplets.len());
        let mut vv: Vec<u64> = Vec::with_capacity(triplets.len());
        let mut rows: Vec<usize> = Vec::with_capacity(triplets.len());
        for &(r, c, v) in triplets.iter() {
            assert!(r < n && c < n, "triplet index out of range");
            if let (Some(&lr), Some(&lc)) = (rows.last(), ci.last()) {
                if lr == r && lc == c {
                    let last = vv.last_mut().unwrap();
                    *last = last.wrapping_add(v);
                    continue;
                }
            }
            rows.push(r);
            ci.push(c);
            vv.push(v);
        }
        let mut kc = Vec::with_capacity(ci.len());
        let mut kv = Vec::with_capacity(vv.len());
        for i in 0..vv.len() {
            if vv[i] != 0 {
                rp[rows[i] + 1] += 1;
                kc.push(ci[i]);
                kv.push(vv[i]);
            }
        }
        for i in 0..n {
            rp[i + 1] += rp[i];
        }
        Self::from_csr_parts_usize(n, &rp, &kc, &kv)
    }

    pub fn from_edges(n: usize, edges: &[(usize, usize)]) -> Self {
        let mut t: Vec<_> = edges.iter().map(|&(a, b)| (a, b, 1u64)).collect();
        Self::from_coo(n, &mut t)
    }

    pub fn from_edges_undirected(n: usize, edges: &[(usize, usize)]) -> Self {
        let mut t = Vec::with_capacity(edges.len() * 2);
        for &(a, b) in edges {
            t.push((a, b, 1u64));
            if a != b {
                t.push((b, a, 1u64));
            }
        }
        Self::from_coo(n, &mut t)
    }

    pub fn from_adjacency<'a>(pairs: impl IntoIterator<Item = (&'a str, &'a str)>) -> (Self, BTreeMap<String, usize>) {
        let mut names: BTreeMap<String, usize> = BTreeMap::new();
        let mut edges = Vec::new();
        for (a, b) in pairs {
            let next = names.len();
            let ai = *names.entry(a.to_string()).or_insert(next);
            let next = names.len();
            let bi = *names.entry(b.to_string()).or_insert(next);
            edges.push((ai, bi));
        }
        (Self::from_edges(names.len(), &edges), names)
    }

    pub fn random(rng: &mut impl Rng, n: usize, m: usize) -> Self {
        assert!(n >= 2, "need at least 2 nodes to avoid self-loops");
        let mut t = Vec::with_capacity(m);
        for _ in 0..m {
            let r = rng.random_range(0..n);
            let mut c = rng.random_range(0..n - 1);
            if c >= r {
                c += 1;
            }
            t.push((r, c, 1u64));
        }
        Self::from_coo(n, &mut t)
    }

    /// N-d Moore-neighbourhood lattice, row-major ids, dimension 0 the least-significant offset digit.
    pub fn lattice(dims: &[usize], torus: bool) -> Self {
        let nd = dims.len();
        let total: usize = dims.iter().product();
        let mut strides = vec![1usize; nd];
        for i in (0..nd.saturating_sub(1)).rev() {
            strides[i] = strides[i + 1] * dims[i + 1];
        }
        let mut t = Vec::new();
        let mut coord = vec![0usize; nd];
        for node in 0..total {
            let mut rem = node;
            for d in 0..nd {
                coord[d] = rem / strides[d];
                rem %= strides[d];
            }
            'offs: for off in 0..3usize.pow(nd as u32) {
                let mut code = off;
                let mut nb = 0usize;
                let mut all_zero = true;
                for d in 0..nd {
                    let delta = (code % 3) as isize - 1;
                    code /= 3;
                    all_zero &= delta == 0;
                    let x = coord[d] as isize + delta;
                    let x = if torus {
                        x.rem_euclid(dims[d] as isize)
                    } else if x < 0 || x >= dims[d] as isize {
                        continue 'offs;
                    } else {
                        x
                    };
                    nb += x as usize * strides[d];
                }
                if !all_zero {
                    t.push((node, nb, 1u64));
                }
            }
        }
        Self::from_coo(total, &mut t)
    }

    /// Symmetric Bernoulli thinning: one draw per stored entry with `r <= c`, mirrored (src/graph_magnus.rs:187-207).
    pub fn thin(&self, rng: &mut impl Rng, density: f64) -> Self {
        let (rp, ci, vv) = self.host();
        let mut t = Vec::new();
        for r in 0..self.n {
            for i in rp[r]..rp[r + 1] {
                let c = ci[i];
                if r <= c && rng.random_range(0.0..1.0) < density {
                    t.push((r, c, vv[i]));
                    if r != c {
                        let back = self.get(c, r);
                        if back != 0 {
                            t.push((c, r, back));
                        }
                    }
                }
            }
        }
        Self::from_coo(self.n, &mut t)
    }

    /// `lattice` built by the engine's device generator (same matrix; 8 M nodes in ~70 ms instead of a host COO sort).
    pub fn lattice_device(dims: &[usize], torus: bool) -> Self {
        let d: Vec<u64> = dims.iter().map(|&x| x as u64).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_lattice(ctx(), d.as_ptr(), d.len() as c_int, torus as c_int, 64, &mut h) });
        Self::wrap(dims.iter().product(), h)
    }

    /// `thin` on the device for the generator the benches use: `StdRng::from_seed(seed)` after `skip` `next_u64` draws
    /// (`bench_repeated_exponentiation` seeds `[42; 32]`, src/graph_magnus.rs:707). Returns the thinned matrix and the
    /// number of draws it consumed, so a caller sharing one generator over several instances can continue it.
    pub fn thin_stdrng(&self, seed: [u8; 32], skip: u64, density: f64) -> (Self, u64) {
        let (mut h, mut taken) = (std::ptr::null_mut(), 0u64);
        check(unsafe { b200_thin(ctx(), self.handle, density, seed.as_ptr(), skip, &mut h, &mut taken) });
        (Self::wrap(self.n, h), taken)
    }

    /// `from_coo` on the device (src/graph_magnus.rs:34-76: sort by (row, column), sum duplicates, drop zeros): the triplets
    /// are uploaded once and ordered by the engine's radix sort.  Plain `+=` duplicate sum, like the reference.
    pub fn from_coo_device(n: usize, triplets: &[(usize, usize, u64)]) -> Self {
        let r: Vec<u32> = triplets.iter().map(|t| t.0 as u32).collect();
        let c: Vec<u32> = triplets.iter().map(|t| t.1 as u32).collect();
        let v: Vec<u64> = triplets.iter().map(|t| t.2).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe {
            b200_csr_from_coo(ctx(), n as u64, n as u64, r.len() as u64, r.as_ptr(), c.as_ptr(), v.as_ptr() as *const c_void, 64, 0, &mut h)
        });
        Self::wrap(n, h)
    }

    /// R-MAT graph of 2^scale nodes generated on the device (BASELINE configs[3]; the reference has no such generator).
    pub fn rmat_device(scale: u32, edge_factor: u64, abc: (f64, f64, f64), seed: u64) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_rmat(ctx(), scale as c_int, edge_factor, abc.0, abc.1, abc.2, seed, 64, &mut h) });
        Self::wrap(1usize << scale, h)
    }

    // ------------------------------------------------------------------ locality pre-pass (src/graph_csr.rs:663-818)
    /// Reorder rows and columns by `perm[new] = old`; returns the reordered matrix (the caller keeps `perm`, as `CsrMatrix.perm`).
    pub fn permute(&self, perm: &[u32]) -> Self {
        assert_eq!(perm.len(), self.n);
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_csr_permute(ctx(), self.handle, perm.as_ptr(), &mut h) });
        Self::wrap(self.n, h)
    }

    /// Reverse Cuthill-McKee: (reordered matrix, perm) with `perm[new] = old`.
    pub fn rcm(&self) -> (Self, Vec<u32>) {
        let mut perm = vec![0u32; self.n];
        check(unsafe { b200_csr_rcm_order(ctx(), self.handle, perm.as_mut_ptr()) });
        (self.permute(&perm), perm)
    }

    /// Undo `perm` (a permute by its inverse).
    pub fn unpermute(&self, perm: &[u32]) -> Self {
        let mut inv = vec![0u32; perm.len()];
        for (new_idx, &old) in perm.iter().enumerate() {
            inv[old as usize] = new_idx as u32;
        }
        self.permute(&inv)
    }

    /// (max |r-c|, mean |r-c|) over the stored entries.
    pub fn bandwidth_stats(&self) -> (usize, f64) {
        let (mut mx, mut avg) = (0u64, 0f64);
        check(unsafe { b200_csr_bandwidth_stats(ctx(), self.handle, &mut mx, &mut avg) });
        (mx as usize, avg)
    }

    // ------------------------------------------------------------------ queries
    pub fn get(&self, r: usize, c: usize) -> u64 {
        let (rp, ci, vv) = self.host();
        match ci[rp[r]..rp[r + 1]].binary_search(&c) {
            Ok(i) => vv[rp[r] + i],
            Err(_) => 0,
        }
    }

    pub fn nnz(&self) -> usize {
        let mut nnz = 0u64;
        check(unsafe { b200_csr_info(self.handle, std::ptr::null_mut(), std::ptr::null_mut(), &mut nnz, std::ptr::null_mut()) });
        nnz as usize
    }

    // ------------------------------------------------------------------ arithmetic: on the GPU
    /// C = self x other (replaces `magnus_spgemm_parallel`, src/graph_magnus.rs:225-232).
    pub fn matmul(&self, other: &Self) -> Self {
        assert_eq!(self.n, other.n);
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_spgemm(ctx(), self.handle, other.handle, &mut h, std::ptr::null_mut()) });
        Self::wrap(self.n, h)
    }

    /// Same engine; kept so call sites written against `MagnusMatrix::matmul_seq` (:235-242) compile unchanged.
    pub fn matmul_seq(&self, other: &Self) -> Self {
        self.matmul(other)
    }

    /// Multiply and return the engine's per-phase measurements.
    pub fn matmul_stats(&self, other: &Self) -> (Self, B200Stats) {
        assert_eq!(self.n, other.n);
        let mut h = std::ptr::null_mut();
        let mut st = B200Stats::default();
        check(unsafe { b200_spgemm(ctx(), self.handle, other.handle, &mut h, &mut st) });
        (Self::wrap(self.n, h), st)
    }

    pub fn add(&self, other: &Self) -> Self {
        assert_eq!(self.n, other.n);
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_csr_add(ctx(), self.handle, other.handle, &mut h) });
        Self::wrap(self.n, h)
    }

    fn same_pattern(&self, other: &Self) -> bool {
        let mut same: c_int = 0;
        check(unsafe { b200_csr_same_pattern(ctx(), self.handle, other.handle, &mut same) });
        same != 0
    }

    pub fn reachability_sum(&self) -> (Self, usize) {
        let mut power = self.clone();
        let mut total = self.clone();
        let mut k = 1;
        loop {
            power = power.matmul(self);
            k += 1;
            let next = total.add(&power);
            if next.nnz() == total.nnz() {
                return (next, k);
            }
            total = next;
        }
    }

    pub fn power_until_stable(&self) -> (Self, usize) {
        let mut current = self.clone();
        let mut k = 0;
        loop {
            let next = current.matmul(&current);
            k += 1;
            if next.nnz() == current.nnz() && next.same_pattern(&current) {
                return (next, k);
            }
            current = next;
        }
    }

    pub fn connected_components(&self) -> Vec<usize> {
        let (closure, _) = self.add(&Self::identity(self.n)).power_until_stable();
        let mut comp = vec![usize::MAX; self.n];
        let mut next = 0;
        for i in 0..self.n {
            if comp[i] != usize::MAX {
                continue;
            }
            comp[i] = next;
            for j in (i + 1)..self.n {
                if closure.get(i, j) > 0 && closure.get(j, i) > 0 {
                    comp[j] = next;
                }
            }
            next += 1;
        }
        comp
    }

    pub fn connected_components_uf(&self) -> Vec<usize> {
        fn find(p: &mut [usize], mut x: usize) -> usize {
            while p[x] != x {
                p[x] = p[p[x]];
                x = p[x];
            }
            x
        }
        let (rp, ci, _) = self.host();
        let mut parent: Vec<usize> = (0..self.n).collect();
        for r in 0..self.n {
            for &c in &ci[rp[r]..rp[r + 1]] {
                let (a, b) = (find(&mut parent, r), find(&mut parent, c));
                if a != b {
                    parent[b] = a;
                }
            }
        }
        let mut ids = BTreeMap::new();
        (0..self.n)
            .map(|i| {
                let root = find(&mut parent, i);
                let next = ids.len();
                *ids.entry(root).or_insert(next)
            })
            .collect()
    }

    pub fn num_components(&self) -> usize {
        self.connected_components_uf().iter().max().map_or(0, |m| m + 1)
    }

    pub fn print(&self) {
        for r in 0..self.n {
            let row: Vec<String> = (0..self.n)
                .map(|c| match self.get(r, c) {
                    0 => ".".to_string(),
                    v => v.to_string(),
                })
                .collect();
            println!("{}", row.join(" "));
        }
    }

    /// Wait for everything queued on the engine's stream (benches call this before stopping a timer).
    pub fn synchronize() {
        check(unsafe { b200_ctx_synchronize(ctx()) });
    }
}

impl NDIndex<u64> for B200Matrix {
    fn ndim(&self) -> usize { 2 }
    fn dim(&self, _axis: usize) -> usize { self.n }
    fn get(&self, ix: &[usize]) -> u64 { B200Matrix::get(self, ix[0], ix[1]) }
    fn set(&mut self, _ix: &[usize], _v: u64) { panic!("B200Matrix is immutable after construction") }
    fn get_opt(&self, ix: &[usize]) -> Option<u64> {
        let (rp, ci, vv) = self.host();
        ci[rp[ix[0]]..rp[ix[0] + 1]].binary_search(&ix[1]).ok().map(|i| vv[rp[ix[0]] + i])
    }
    fn is_sparse_2d(&self) -> bool { true }
    fn sparse_row_nnz(&self, row: usize) -> usize {
        let rp = &self.host().0;
        rp[row + 1] - rp[row]
    }
    fn sparse_row_entry(&self, row: usize, idx: usize) -> (usize, u64) {
        let (rp, ci, vv) = self.host();
        (ci[rp[row] + idx], vv[rp[row] + idx])
    }
}


/// One rank (one GPU) of a multi-GPU job: `b200_comm` bound to this process's engine context.
/// One process per GPU: rank 0 calls `B200Comm::unique_id()`, the launcher carries the 128 bytes to the other ranks
/// (MPI, a file, a socket), every rank calls `B200Comm::init_rank`.  The path shards by rows of the left operand:
/// `broadcast` replicates the right operand once, `shard` gives every rank its product-balanced row block, and the
/// power chain `A^k = A^(k-1) x A` then runs on the blocks without communication; `allgather` assembles the blocks
/// (the optional gather of C, or the per-step exchange of a squaring chain, src/graph_csr.rs:561-575).
pub struct B200Comm {
    raw: *mut B200CommRaw,
    pub rank: usize,
    pub size: usize,
}

impl Drop for B200Comm {
    fn drop(&mut self) {
        unsafe { b200_comm_destroy(self.raw) };
    }
}

impl B200Comm {
    pub fn unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        check(unsafe { b200_comm_unique_id(id.as_mut_ptr()) });
        id
    }

    pub fn init_rank(nranks: usize, rank: usize, id: &[u8; 128]) -> Self {
        let mut raw = std::ptr::null_mut();
        check(unsafe { b200_comm_init_rank(ctx(), nranks as c_int, rank as c_int, id.as_ptr(), &mut raw) });
        Self { raw, rank, size: nranks }
    }

    /// Replicate `src` (given on `root`, `None` elsewhere) on every rank.
    pub fn broadcast(&self, src: Option<&B200Matrix>, root: usize) -> B200Matrix {
        let mut h = std::ptr::null_mut();
        let p = src.map_or(std::ptr::null(), |m| m.handle as *const B200Csr);
        check(unsafe { b200_comm_broadcast_csr(self.raw, p, root as c_int, &mut h) });
        let (mut rows, mut cols, mut nnz, mut bits) = (0u64, 0u64, 0u64, 0);
        check(unsafe { b200_csr_info(h, &mut rows, &mut cols, &mut nnz, &mut bits) });
        B200Matrix::wrap(rows as usize, h)
    }

    /// This rank's rows of `a` for `a x b`, cut so that every rank holds the same share of the intermediate products.
    pub fn shard(&self, a: &B200Matrix, b: &B200Matrix) -> B200Matrix {
        let mut cuts = vec![0u64; self.size + 1];
        check(unsafe { b200_shard_rows_by_products(ctx(), a.handle, b.handle, self.size as c_int, cuts.as_mut_ptr()) });
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_csr_row_block(ctx(), a.handle, cuts[self.rank], cuts[self.rank + 1], &mut h) });
        B200Matrix::wrap((cuts[self.rank + 1] - cuts[self.rank]) as usize, h)
    }

    /// Row blocks in rank order -> the whole matrix on every rank.
    pub fn allgather(&self, block: &B200Matrix) -> B200Matrix {
        let mut h = std::ptr::null_mut();
        check(unsafe { b200_comm_allgather_csr(self.raw, block.handle, &mut h) });
        let (mut rows, mut cols, mut nnz, mut bits) = (0u64, 0u64, 0u64, 0);
        check(unsafe { b200_csr_info(h, &mut rows, &mut cols, &mut nnz, &mut bits) });
        B200Matrix::wrap(rows as usize, h)
    }

    /// Max over the ranks (e.g. of a step time in ms).
    pub fn max_f64(&self, v: f64) -> f64 {
        let mut x = v;
        check(unsafe { b200_comm_allreduce(self.raw, &mut x as *mut f64 as *mut c_void, 1, 3) });
        x
    }

    /// Sum over the ranks (e.g. of intermediate products).
    pub fn sum_u64(&self, v: u64) -> u64 {
        let mut x = v;
        check(unsafe { b200_comm_allreduce(self.raw, &mut x as *mut u64 as *mut c_void, 1, 0) });
        x
    }
}
