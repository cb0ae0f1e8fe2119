//! `RIFull` / `MatrixFull` / `MatrixUpper` with the reference's field names and layouts (src/ri.rs:18-24,
//! src/matrix/mod.rs:472-480, src/matrix/matrixupper.rs:231-234) and the hot-path methods bound to librest_b200.
//! A maintainer of rest_tensors keeps the crate's own structs and pastes the method bodies; this module exists so that
//! the FFI crate is usable (and testable) on its own.  Column-major everywhere, public fields like the reference.
use crate::ffi::*;
use std::ffi::c_int;

#[derive(Clone, Debug, PartialEq)]
pub struct MatrixFull { pub size: [usize; 2], pub indicing: [usize; 2], pub data: Vec<f64> }
#[derive(Clone, Debug, PartialEq)]
pub struct MatrixUpper { pub size: usize, pub data: Vec<f64> }
#[derive(Clone, Debug, PartialEq)]
pub struct RIFull { pub size: [usize; 3], pub indicing: [usize; 3], pub data: Vec<f64> }

fn ci(v: usize) -> c_int { c_int::try_from(v).expect("dimension does not fit the 32-bit FFI integer (src/external_libs/mod.rs:18-20)") }

impl MatrixFull {
    /// src/matrix/matrixfull.rs:190-205
    pub fn new(size: [usize; 2], v: f64) -> MatrixFull {
        MatrixFull { size, indicing: [1, size[0]], data: vec![v; size[0] * size[1]] }
    }
    /// src/matrix/matrixfull.rs:227-240: panics when the vector is too short, keeps a longer one
    pub fn from_vec(size: [usize; 2], data: Vec<f64>) -> MatrixFull {
        let len = size[0] * size[1];
        if len > data.len() {
            panic!("Error: inconsistency happens when formating a matrix from a given vector, (length from size, length of new vector) = ({},{})", len, data.len());
        }
        MatrixFull { size, indicing: [1, size[0]], data }
    }
    /// src/matrix/matrixfull.rs:638-646
    pub fn to_matrixupper(&self) -> MatrixUpper {
        if self.size[0] != self.size[1] { panic!("Error: Nonsymmetric matrix cannot be converted to the upper format"); }
        let n = self.size[0];
        let mut out = vec![0.0f64; n * (n + 1) / 2];
        unsafe { check(rb_host_to_matrixupper(self.data.as_ptr(), n as i64, out.as_mut_ptr()), "to_matrixupper"); }
        MatrixUpper { size: out.len(), data: out }
    }
    /// src/matrix/matrixfull.rs:579-614
    pub fn transpose(&self) -> MatrixFull {
        let mut out = MatrixFull::new([self.size[1], self.size[0]], 0.0);
        unsafe { check(rb_host_matrix_transpose(self.data.as_ptr(), self.size[0] as i64, self.size[1] as i64, out.data.as_mut_ptr()), "transpose"); }
        out
    }
    /// src/matrix/mod.rs:545-648: c += p * b (unfused multiply-add, bit-identical to the Rust loop)
    pub fn self_scaled_add(&mut self, p: &MatrixFull, b: f64) {
        if self.size != p.size { panic!("Error: the two matrices have different shapes"); }
        unsafe { check(rb_host_axpy(0, self.data.as_mut_ptr(), p.data.as_ptr(), 0.0, b, self.data.len() as i64), "self_scaled_add"); }
    }
    /// src/matrix/matrix_blas_lapack.rs:180-252 (`_dgemm_full`): C = alpha op(A) op(B) + beta C on whole matrices
    pub fn dgemm_full(a: &MatrixFull, opa: char, b: &MatrixFull, opb: char, c: &mut MatrixFull, alpha: f64, beta: f64) {
        let (m, k) = if opa == 'N' { (a.size[0], a.size[1]) } else { (a.size[1], a.size[0]) };
        let (kb, n) = if opb == 'N' { (b.size[0], b.size[1]) } else { (b.size[1], b.size[0]) };
        if k != kb || c.size != [m, n] { panic!("Error: Inconsistency happens to perform dgemm w.r.t. op(a)*op(b) -> c"); }
        unsafe {
            check(rb_host_dgemm(opa as u8 as _, opb as u8 as _, ci(m), ci(n), ci(k), alpha, a.data.as_ptr(), ci(a.size[0].max(1)),
                                b.data.as_ptr(), ci(b.size[0].max(1)), beta, c.data.as_mut_ptr(), ci(m.max(1))), "_dgemm_full");
        }
    }
    /// src/matrix/matrix_blas_lapack.rs:392-413 (`_dsyrk`): only the `uplo` triangle of C is touched
    pub fn dsyrk(a: &MatrixFull, c: &mut MatrixFull, uplo: char, trans: char, alpha: f64, beta: f64) {
        if c.size[0] != c.size[1] { panic!("Error: the matrix C is not square for dsyrk"); }
        let n = c.size[0];
        let k = if trans == 'N' || trans == 'n' { a.size[1] } else { a.size[0] };
        let lda = if trans == 'N' || trans == 'n' { n } else { k };
        unsafe {
            check(rb_host_dsyrk(uplo as u8 as _, trans as u8 as _, ci(n), ci(k), alpha, a.data.as_ptr(), ci(lda.max(1)), beta,
                                c.data.as_mut_ptr(), ci(n.max(1))), "_dsyrk");
        }
    }
}

impl MatrixUpper {
    pub fn new(size: usize, v: f64) -> MatrixUpper { MatrixUpper { size, data: vec![v; size] } }
    /// src/matrix/matrixupper.rs:330-373: `None` unless the length is triangular; empty for length 0
    pub fn to_matrixfull(&self) -> Option<MatrixFull> {
        let len = self.data.len();
        let n = (((1.0 + 8.0 * len as f64).sqrt() * 0.5) - 0.5) as usize;
        if n * (n + 1) / 2 != len { return None; }
        let mut out = MatrixFull::new([n, n], 0.0);
        if len > 0 {
            unsafe { check(rb_host_to_matrixfull(self.data.as_ptr(), len as i64, out.data.as_mut_ptr()), "to_matrixfull"); }
        }
        Some(out)
    }
}

impl RIFull {
    /// src/ri.rs:27-40
    pub fn new(size: [usize; 3], v: f64) -> RIFull {
        RIFull { size, indicing: [1, size[0], size[0] * size[1]], data: vec![v; size[0] * size[1] * size[2]] }
    }
    /// src/ri.rs:53-70
    pub fn from_vec(size: [usize; 3], data: Vec<f64>) -> RIFull {
        let len = size[0] * size[1] * size[2];
        if len > data.len() {
            panic!("Error: inconsistency happens when formating a tensor from a given vector, (length from size, length of new vector) = ({},{})", len, data.len());
        }
        RIFull { size, indicing: [1, size[0], size[0] * size[1]], data }
    }
    /// src/ri.rs:356-408 (`ao2mo` == `ao2mo_v02`): ri3mo[P, a, b], P fastest, through the Fortran-ABI symbol
    pub fn ao2mo(&self, eigenvector: &MatrixFull) -> RIFull {
        let (nb, ns, nx) = (eigenvector.size[0], eigenvector.size[1], self.size[2]);
        let mut out = RIFull::new([nx, ns, ns], 0.0);
        let (ns_i, nb_i, nx_i) = (ci(ns), ci(nb), ci(nx));
        unsafe { compat::ri_ao2mo_f_(eigenvector.data.as_ptr(), self.data.as_ptr(), out.data.as_mut_ptr(), &ns_i, &nb_i, &nx_i); }
        out
    }
    /// north-star occ-vir form: out[P, a, b] = sum C_L[mu, a] A[mu, nu, P] C_R[nu, b]
    pub fn ao2mo_rect(&self, c_left: &MatrixFull, c_right: &MatrixFull) -> RIFull {
        let (nb, nl, nr, nx) = (self.size[0], c_left.size[1], c_right.size[1], self.size[2]);
        let mut out = RIFull::new([nx, nl, nr], 0.0);
        unsafe {
            check(rb_host_ri_ao2mo(c_left.data.as_ptr(), ci(nl), c_right.data.as_ptr(), ci(nr), self.data.as_ptr(),
                                   out.data.as_mut_ptr(), ci(nb), ci(nx)), "ao2mo_rect");
        }
        out
    }
    /// d_P = sum_{mu nu} ri3ao[mu, nu, P] D[mu, nu]   (REST composes this from `_dgemv`)
    pub fn ri_dp(&self, dm: &MatrixFull) -> Vec<f64> {
        let mut d = vec![0.0f64; self.size[2]];
        unsafe { check(rb_host_ri_dp(self.data.as_ptr(), dm.data.as_ptr(), d.as_mut_ptr(), ci(self.size[0]), ci(self.size[2])), "ri_dp"); }
        d
    }
    /// J = sum_P ri3ao[:, :, P] d_P
    pub fn ri_j(&self, d: &[f64]) -> MatrixFull {
        let nb = self.size[0];
        let mut j = MatrixFull::new([nb, nb], 0.0);
        unsafe { check(rb_host_ri_j(self.data.as_ptr(), d.as_ptr(), j.data.as_mut_ptr(), ci(nb), ci(self.size[2])), "ri_j"); }
        j
    }
    /// K = sum_P (A_P C~)(A_P C~)^T, C~ = C_occ diag(sqrt(n_occ))
    pub fn ri_k(&self, ct: &MatrixFull) -> MatrixFull {
        let nb = self.size[0];
        let mut k = MatrixFull::new([nb, nb], 0.0);
        unsafe { check(rb_host_ri_k(self.data.as_ptr(), ct.data.as_ptr(), ci(ct.size[1]), k.data.as_mut_ptr(), ci(nb), ci(self.size[2])), "ri_k"); }
        k
    }
    /// ao2mo + d_P + J + K in ONE streaming pass over the host tensor (every slab crosses PCIe once)
    pub fn ao2mo_jk(&self, c: &MatrixFull, dm: &MatrixFull, ct: &MatrixFull) -> (RIFull, Vec<f64>, MatrixFull, MatrixFull) {
        let (nb, ns, nx) = (self.size[0], c.size[1], self.size[2]);
        let mut mo = RIFull::new([nx, ns, ns], 0.0);
        let mut d = vec![0.0f64; nx];
        let (mut j, mut k) = (MatrixFull::new([nb, nb], 0.0), MatrixFull::new([nb, nb], 0.0));
        unsafe {
            check(rb_host_ri_ao2mo_jk(c.data.as_ptr(), ci(ns), c.data.as_ptr(), ci(ns), self.data.as_ptr(), mo.data.as_mut_ptr(),
                                      ci(nb), ci(nx), dm.data.as_ptr(), ct.data.as_ptr(), ci(ct.size[1]), d.as_mut_ptr(),
                                      j.data.as_mut_ptr(), k.data.as_mut_ptr()), "ao2mo_jk");
        }
        (mo, d, j, k)
    }
    /// ao2mo_jk shipping only the a <= b part of ri3mo: upper[P + nx * (b(b+1)/2 + a)] (MatrixUpper's pair index per P); for the
    /// symmetric slabs of an RI tensor that is the whole result at half the device -> host traffic
    /// `symmetric_slabs`: the caller guarantees self[mu, nu, P] == self[nu, mu, P]; only mu <= nu is uploaded then
    pub fn ao2mo_jk_upper(&self, c: &MatrixFull, dm: &MatrixFull, ct: &MatrixFull, symmetric_slabs: bool) -> (Vec<f64>, Vec<f64>, MatrixFull, MatrixFull) {
        let (nb, ns, nx) = (self.size[0], c.size[1], self.size[2]);
        let mut upper = vec![0.0f64; nx * (ns * (ns + 1) / 2)];
        let mut d = vec![0.0f64; nx];
        let (mut j, mut k) = (MatrixFull::new([nb, nb], 0.0), MatrixFull::new([nb, nb], 0.0));
        unsafe {
            let st = if symmetric_slabs {
                rb_host_ri_ao2mo_jk_symm(c.data.as_ptr(), ci(ns), self.data.as_ptr(), upper.as_mut_ptr(), ci(nb), ci(nx),
                                         dm.data.as_ptr(), ct.data.as_ptr(), ci(ct.size[1]), d.as_mut_ptr(),
                                         j.data.as_mut_ptr(), k.data.as_mut_ptr())
            } else {
                rb_host_ri_ao2mo_jk_upper(c.data.as_ptr(), ci(ns), self.data.as_ptr(), upper.as_mut_ptr(), ci(nb), ci(nx),
                                          dm.data.as_ptr(), ct.data.as_ptr(), ci(ct.size[1]), d.as_mut_ptr(),
                                          j.data.as_mut_ptr(), k.data.as_mut_ptr())
            };
            check(st, "ao2mo_jk_upper");
        }
        (upper, d, j, k)
    }
    /// src/ri.rs:227-294; which = 0 jik, 1 jki, 2 kji, 3 ikj
    pub fn transpose(&self, which: usize) -> RIFull {
        let [i, j, k] = self.size;
        let size = match which { 0 => [j, i, k], 1 => [j, k, i], 2 => [k, j, i], 3 => [i, k, j], _ => panic!("transpose: which must be 0..3") };
        let mut out = RIFull::new(size, 0.0);
        unsafe { check(rb_host_ri_transpose(self.data.as_ptr(), i as i64, j as i64, k as i64, which as c_int, out.data.as_mut_ptr()), "transpose"); }
        out
    }
    /// src/ri.rs:297-306
    pub fn rifull_to_matfull_symm(&self) -> MatrixFull {
        let (nb, nx) = (self.size[0], self.size[2]);
        let mut out = MatrixFull::new([nb * (nb + 1) / 2, nx], 0.0);
        unsafe { check(rb_host_ri_pack_symm(self.data.as_ptr(), nb as i64, nx as i64, out.data.as_mut_ptr()), "rifull_to_matfull_symm"); }
        out
    }
    /// src/ri.rs:410-419 -> src/external_libs/mod.rs:168-190 -> copy_rr_
    pub fn copy_from_ri(&mut self, rx: std::ops::Range<usize>, ry: std::ops::Range<usize>, rz: std::ops::Range<usize>, from: &RIFull,
                        fx: std::ops::Range<usize>, fy: std::ops::Range<usize>, fz: std::ops::Range<usize>) {
        if rx.len() != fx.len() || ry.len() != fy.len() || rz.len() != fz.len() {
            panic!("Error: the data block for copy has different size between ri 3D-tensors");
        }
        let v = [ci(rx.len()), ci(ry.len()), ci(rz.len())];
        let f = [ci(from.size[0]), ci(from.size[1]), ci(from.size[2]), ci(fx.start), ci(fy.start), ci(fz.start)];
        let t = [ci(self.size[0]), ci(self.size[1]), ci(self.size[2]), ci(rx.start), ci(ry.start), ci(rz.start)];
        unsafe {
            compat::copy_rr_(&v[0], &v[1], &v[2], from.data.as_ptr(), &f[0], &f[1], &f[2], &f[3], &f[4], &f[5], self.data.as_mut_ptr(),
                             &t[0], &t[1], &t[2], &t[3], &t[4], &t[5]);
        }
    }
}
