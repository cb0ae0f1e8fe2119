// rest_tensors.hpp -- header-only C++ mirror of the reference's host API for the RI hot path, on top of the C ABI
// (include/rest_b200.h).  The reference is a compiled (Rust) crate and no Rust toolchain exists in this image, so this
// is the compiled-language host side: same type names, public fields (size / indicing / data), method names, argument
// meaning and error behaviour (a Rust panic! is a std::runtime_error here, Option::None is std::nullopt).
//   RIFull        reference src/ri.rs:18-433
//   MatrixFull    reference src/matrix/mod.rs:472-480, src/matrix/matrixfull.rs
//   MatrixUpper   reference src/matrix/matrixupper.rs:231-420, src/index.rs:209-233
//   _dgemm_full, _dsyrk, _dgemv, _dsymm   reference src/matrix/matrix_blas_lapack.rs
// Every numerical operation and every bulk data movement is a call into librest_b200.so; std::vector only owns buffers.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rest_b200.h"

namespace rest_tensors {

using Range = std::pair<size_t, size_t>; // half-open start..end

inline void rb_check(int status, const char *what)
{
    if (status != RB_OK) throw std::runtime_error(std::string(what) + " failed: " + rb_last_error());
}
inline size_t rlen(const Range &r) { return r.second > r.first ? r.second - r.first : 0; }

struct MatrixUpper;
struct RIFull;

struct MatrixFull {
    std::array<size_t, 2> size{0, 0};
    std::array<size_t, 2> indicing{0, 0};
    std::vector<double> data;

    static MatrixFull make(std::array<size_t, 2> size, double v)
    {
        MatrixFull m;
        m.size = size;
        m.indicing = {1, size[0]};
        m.data.assign(size[0] * size[1], v);
        return m;
    }
    static MatrixFull empty() { return MatrixFull{}; }
    // matrixfull.rs:227-240: panics when the vector is too short, keeps a surplus
    static MatrixFull from_vec(std::array<size_t, 2> size, std::vector<double> v)
    {
        if (size[0] * size[1] > v.size())
            throw std::runtime_error("Error: inconsistency happens when formating a matrix from a given vector");
        MatrixFull m;
        m.size = size;
        m.indicing = {1, size[0]};
        m.data = std::move(v);
        return m;
    }
    bool check_shape(const MatrixFull &o) const { return size == o.size; }
    // matrixfull.rs:242-253: same data, new shape (the element count must agree)
    void reshape(std::array<size_t, 2> new_size)
    {
        if (new_size[0] * new_size[1] != data.size()) throw std::runtime_error("Error: the reshaped matrix has a different number of elements");
        size = new_size;
        indicing = {1, new_size[0]};
    }
    inline RIFull to_rifull(size_t i, size_t j, size_t k) const; // matrixfull.rs:674-686
    // matrixfull.rs:1398-1406 / matrix_blas_lapack.rs:739-774: self = alpha op(a) op(b) + beta self, NO shape check (like the reference)
    void lapack_dgemm(const MatrixFull &a, const MatrixFull &b, char opa, char opb, double alpha, double beta)
    {
        const size_t m = size[0], n = size[1], k = (opa == 'N' || opa == 'n') ? a.size[1] : a.size[0];
        rb_check(rb_host_dgemm(opa, opb, (int)m, (int)n, (int)k, alpha, a.data.data(), (int)std::max<size_t>(a.size[0], 1),
                               b.data.data(), (int)std::max<size_t>(b.size[0], 1), beta, data.data(), (int)std::max<size_t>(m, 1)),
                 "lapack_dgemm");
    }
    // matrix_blas_lapack.rs:714-730 (MatrixFullSlice::ddot): a * b as a new matrix, None on an inner-dimension mismatch
    std::optional<MatrixFull> ddot(const MatrixFull &b) const
    {
        if (size[1] != b.size[0]) return std::nullopt;
        MatrixFull c = make({size[0], b.size[1]}, 0.0);
        c.lapack_dgemm(*this, b, 'N', 'N', 1.0, 1.0); // beta = 1 on a zeroed C, as the reference does
        return c;
    }

    MatrixFull transpose() const // matrixfull.rs:579-596
    {
        MatrixFull t = make({size[1], size[0]}, 0.0);
        if (!data.empty())
            rb_check(rb_host_matrix_transpose(data.data(), (int64_t)size[0], (int64_t)size[1], t.data.data()), "transpose");
        return t;
    }
    inline MatrixUpper to_matrixupper() const; // matrixfull.rs:638-646

    // matrix_blas_lapack.rs:775-797: (eigenvectors, eigenvalues ascending); the lower triangle is read
    std::pair<MatrixFull, std::vector<double>> lapack_dsyev() const
    {
        if (size[0] != size[1]) throw std::runtime_error("Error in _dsyev: the algorithm is only vaild for real symmetric matrices");
        MatrixFull z = make(size, 0.0);
        std::vector<double> w(size[0], 0.0);
        rb_check(rb_host_dsyev('V', (int)size[0], data.data(), w.data(), z.data.data()), "lapack_dsyev");
        return {std::move(z), std::move(w)};
    }
    // matrix_blas_lapack.rs:599-652 / 1004-1062: A^p over the eigenvalues >= threshold
    MatrixFull lapack_power(double p, double threshold) const
    {
        if (size[0] != size[1]) throw std::runtime_error("Error: The matrix for power operations should be NxN");
        MatrixFull om = make(size, 0.0);
        int kept = 0;
        rb_check(rb_host_power((int)size[0], data.data(), p, threshold, om.data.data(), &kept), "lapack_power");
        return om;
    }

    // matrixfull.rs:1388-1396 -> copy_mm_
    void copy_from_matr(Range rx, Range ry, const MatrixFull &from, Range frx, Range fry)
    {
        if (rlen(rx) != rlen(frx) || rlen(ry) != rlen(fry))
            throw std::runtime_error("Error: the data block for copy has different size between two matrices");
        int xl = (int)rlen(rx), yl = (int)rlen(ry), fx = (int)from.size[0], fy = (int)from.size[1], fxs = (int)frx.first,
            fys = (int)fry.first, tx = (int)size[0], ty = (int)size[1], txs = (int)rx.first, tys = (int)ry.first;
        copy_mm_(&xl, &yl, from.data.data(), &fx, &fy, &fxs, &fys, data.data(), &tx, &ty, &txs, &tys);
    }
    // matrix/mod.rs:545-648
    void self_scaled_add(const MatrixFull &bm, double b) { need(bm); rb_check(rb_host_axpy(0, data.data(), bm.data.data(), 0, b, (int64_t)data.size()), "self_scaled_add"); }
    void self_general_add(const MatrixFull &bm, double a, double b) { need(bm); rb_check(rb_host_axpy(1, data.data(), bm.data.data(), a, b, (int64_t)data.size()), "self_general_add"); }
    void self_multiple(double a) { rb_check(rb_host_axpy(2, data.data(), nullptr, a, 0, (int64_t)data.size()), "self_multiple"); }
    void self_add(const MatrixFull &bm) { need(bm); rb_check(rb_host_axpy(3, data.data(), bm.data.data(), 0, 0, (int64_t)data.size()), "self_add"); }
    void self_sub(const MatrixFull &bm) { need(bm); rb_check(rb_host_axpy(4, data.data(), bm.data.data(), 0, 0, (int64_t)data.size()), "self_sub"); }

  private:
    void need(const MatrixFull &bm) const
    {
        if (!check_shape(bm)) throw std::runtime_error("Error: Shape inconsistency happens when plus two matrices");
    }
};

struct MatrixUpper {
    size_t size = 0; // packed length n(n+1)/2
    std::vector<double> data;

    static MatrixUpper from_vec(size_t size, std::vector<double> v)
    {
        if (size > v.size()) throw std::runtime_error("Error: inconsistency happens when formating a matrix from a given vector");
        MatrixUpper m;
        m.size = size;
        m.data = std::move(v);
        return m;
    }
    // index.rs:209-226
    std::optional<size_t> index2d(size_t i, size_t j) const
    {
        if (i > j) std::swap(i, j);
        size_t tp = (j + 1) * j / 2 + i;
        return tp < data.size() ? std::optional<size_t>(tp) : std::nullopt;
    }
    size_t dim() const { return (size_t)(std::sqrt(1.0 + 8.0 * (double)size) * 0.5 - 0.5); }
    // matrix_blas_lapack.rs:1075-1095: (eigenvectors [n, n], eigenvalues ascending)
    std::pair<MatrixFull, std::vector<double>> lapack_dspevx() const
    {
        const size_t n = dim();
        if (n * (n + 1) / 2 != size) throw std::runtime_error("lapack_dspevx: the packed length is not triangular");
        MatrixFull z = MatrixFull::make({n, n}, 0.0);
        std::vector<double> w(n, 0.0);
        int found = 0;
        rb_check(rb_host_dspevx((int)n, data.data(), w.data(), z.data.data(), &found), "lapack_dspevx");
        return {std::move(z), std::move(w)};
    }
    // matrix_blas_lapack.rs:1096-1147: A x = lambda B x, the num_orb lowest pairs; eigenvectors [n, num_orb]
    std::pair<MatrixFull, std::vector<double>> lapack_dspgvx(const MatrixUpper &ovlp, size_t num_orb) const
    {
        const size_t n = dim();
        if (ovlp.size != size) throw std::runtime_error("ERROR:: _dspgvx for BasicMatUp, Matr_A and Matr_B have different size");
        if (num_orb > n) throw std::runtime_error("Error:: The number of outcoming eigenvectors is unequal to the orbital number");
        MatrixFull z = MatrixFull::make({n, num_orb}, 0.0);
        std::vector<double> w(num_orb, 0.0);
        rb_check(rb_host_dspgvx((int)n, data.data(), ovlp.data.data(), (int)num_orb, w.data(), z.data.data()), "lapack_dspgvx");
        return {std::move(z), std::move(w)};
    }
    // matrixupper.rs:330-373
    std::optional<MatrixFull> to_matrixfull() const
    {
        if (data.empty()) return MatrixFull::empty();
        size_t n = (size_t)(std::sqrt(1.0 + 8.0 * (double)size) * 0.5 - 0.5);
        if (n * (n + 1) / 2 != size) return std::nullopt;
        MatrixFull f = MatrixFull::make({n, n}, 0.0);
        rb_check(rb_host_to_matrixfull(data.data(), (int64_t)size, f.data.data()), "to_matrixfull");
        return f;
    }
};

inline MatrixUpper MatrixFull::to_matrixupper() const
{
    if (size[0] != size[1]) throw std::runtime_error("Error: Nonsymmetric matrix cannot be converted to the upper format");
    size_t n = size[0];
    MatrixUpper u;
    u.size = n * (n + 1) / 2;
    u.data.assign(u.size, 0.0);
    if (n) rb_check(rb_host_to_matrixupper(data.data(), (int64_t)n, u.data.data()), "to_matrixupper");
    return u;
}

struct RIFull {
    std::array<size_t, 3> size{0, 0, 0};
    std::array<size_t, 3> indicing{0, 0, 0};
    std::vector<double> data;

    static RIFull make(std::array<size_t, 3> size, double v)
    {
        RIFull r;
        r.size = size;
        r.indicing = {1, size[0], size[0] * size[1]};
        r.data.assign(size[0] * size[1] * size[2], v);
        return r;
    }
    static RIFull from_vec(std::array<size_t, 3> size, std::vector<double> v) // ri.rs:57-70
    {
        if (size[0] * size[1] * size[2] > v.size())
            throw std::runtime_error("Error: inconsistency happens when formating a tensor from a given vector");
        RIFull r;
        r.size = size;
        r.indicing = {1, size[0], size[0] * size[1]};
        r.data = std::move(v);
        return r;
    }
    bool check_shape(const RIFull &o) const { return size == o.size; }
    // ri.rs:92-100: pointer to slab P (zero-copy view)
    const double *get_reducing_matrix(size_t p) const { return data.data() + indicing[2] * p; }
    // ri.rs:190-198: [begin, end) pointers of the slab range -- the P-shard primitive
    std::pair<const double *, const double *> iter_auxbas(Range r) const
    {
        size_t chunk = size[0] * size[1];
        return {data.data() + chunk * r.first, data.data() + chunk * r.second};
    }
    // ri.rs:102-115: columns `cols` of slab p (start pointer, element count) -- a zero-copy view
    std::pair<const double *, size_t> get_reducing_matrix_columns(Range cols, size_t p) const
    {
        return {data.data() + p * indicing[2] + size[0] * cols.first, size[0] * rlen(cols)};
    }
    // ri.rs:117-128: the flattened x-runs of a sub-box, z outer / y inner: (pointer, length) per run, zero copy
    std::vector<std::pair<const double *, size_t>> get_slices(Range x, Range y, Range z) const
    {
        std::vector<std::pair<const double *, size_t>> runs;
        runs.reserve(rlen(y) * rlen(z));
        for (size_t zz = z.first; zz < z.second; ++zz)
            for (size_t yy = y.first; yy < y.second; ++yy)
                runs.emplace_back(data.data() + x.first + yy * indicing[1] + zz * indicing[2], rlen(x));
        return runs;
    }
    // ri.rs:309-326: metadata-only reshapes (the reference clones the data; so does this)
    MatrixFull rifull_to_matfull_ij_k() const { return MatrixFull::from_vec({size[0] * size[1], size[2]}, data); }
    MatrixFull rifull_to_matfull_i_jk() const { return MatrixFull::from_vec({size[0], size[1] * size[2]}, data); }
    // north-star occ-vir form: out[P, a, b] = sum C_L[mu, a] A[mu, nu, P] C_R[nu, b]
    RIFull ao2mo_rect(const MatrixFull &c_left, const MatrixFull &c_right) const
    {
        RIFull mo = make({size[2], c_left.size[1], c_right.size[1]}, 0.0);
        rb_check(rb_host_ri_ao2mo(c_left.data.data(), (int)c_left.size[1], c_right.data.data(), (int)c_right.size[1], data.data(),
                                  mo.data.data(), (int)size[0], (int)size[2]), "ao2mo_rect");
        return mo;
    }
    // One streaming pass: ao2mo + d_P + J + K with every P-chunk uploaded once (not in the reference, which makes the calls one by
    // one).  `upper`: ri3mo comes back as its a <= b pairs, upper[P + nx * (b (b + 1) / 2 + a)]; `symmetric_slabs`: the caller
    // guarantees (*this)[mu, nu, P] == (*this)[nu, mu, P] and only mu <= nu is uploaded.
    struct StepResult { std::vector<double> ri3mo; std::vector<double> d; MatrixFull j, k; };
    StepResult ao2mo_jk(const MatrixFull &c, const MatrixFull &dm, const MatrixFull &ct, bool upper = false,
                        bool symmetric_slabs = false) const
    {
        const int nb = (int)c.size[0], ns = (int)c.size[1], nx = (int)size[2], no = (int)ct.size[1];
        StepResult r{std::vector<double>((size_t)nx * (upper ? (size_t)ns * (ns + 1) / 2 : (size_t)ns * ns), 0.0),
                     std::vector<double>((size_t)nx, 0.0), MatrixFull::make({(size_t)nb, (size_t)nb}, 0.0),
                     MatrixFull::make({(size_t)nb, (size_t)nb}, 0.0)};
        if (!upper)
            rb_check(rb_host_ri_ao2mo_jk(c.data.data(), ns, c.data.data(), ns, data.data(), r.ri3mo.data(), nb, nx, dm.data.data(),
                                         ct.data.data(), no, r.d.data(), r.j.data.data(), r.k.data.data()), "ao2mo_jk");
        else
            rb_check((symmetric_slabs ? rb_host_ri_ao2mo_jk_symm : rb_host_ri_ao2mo_jk_upper)(
                         c.data.data(), ns, data.data(), r.ri3mo.data(), nb, nx, dm.data.data(), ct.data.data(), no, r.d.data(),
                         r.j.data.data(), r.k.data.data()), "ao2mo_jk(upper)");
        return r;
    }
    // ri.rs:356-408
    RIFull ao2mo(const MatrixFull &eigenvector) const { return ao2mo_v02(eigenvector); }
    RIFull ao2mo_v02(const MatrixFull &eigenvector) const
    {
        int nb = (int)eigenvector.size[0], ns = (int)eigenvector.size[1], nx = (int)size[2];
        RIFull mo = make({(size_t)nx, (size_t)ns, (size_t)ns}, 0.0);
        ri_ao2mo_f_(eigenvector.data.data(), data.data(), mo.data.data(), &ns, &nb, &nx);
        return mo;
    }
    // ri.rs:227-294
    RIFull transpose_jik() const { return tr(0, {size[1], size[0], size[2]}); }
    RIFull transpose_jki() const { return tr(1, {size[1], size[2], size[0]}); }
    RIFull transpose_kji() const { return tr(2, {size[2], size[1], size[0]}); }
    RIFull transpose_ikj() const { return tr(3, {size[0], size[2], size[1]}); }
    // ri.rs:297-306
    MatrixFull rifull_to_matfull_symm() const
    {
        size_t nao = size[0], naux = size[2];
        MatrixFull out = MatrixFull::make({nao * (nao + 1) / 2, naux}, 0.0);
        if (nao && naux) rb_check(rb_host_ri_pack_symm(data.data(), (int64_t)nao, (int64_t)naux, out.data.data()), "rifull_to_matfull_symm");
        return out;
    }
    // ri.rs:345-354
    void self_scaled_add(const RIFull &bm, double b)
    {
        if (!check_shape(bm)) throw std::runtime_error("Error: Shape inconsistency happens when plus two matrices");
        rb_check(rb_host_axpy(0, data.data(), bm.data.data(), 0, b, (int64_t)data.size()), "self_scaled_add");
    }
    // ri.rs:410-419 -> copy_rr_
    void copy_from_ri(Range rx, Range ry, Range rz, const RIFull &from, Range fx, Range fy, Range fz)
    {
        if (rlen(rx) != rlen(fx) || rlen(ry) != rlen(fy) || rlen(rz) != rlen(fz))
            throw std::runtime_error("Error: the data block for copy has different size between ri 3D-tensors");
        int xl = (int)rlen(rx), yl = (int)rlen(ry), zl = (int)rlen(rz);
        int f[6] = {(int)from.size[0], (int)from.size[1], (int)from.size[2], (int)fx.first, (int)fy.first, (int)fz.first};
        int t[6] = {(int)size[0], (int)size[1], (int)size[2], (int)rx.first, (int)ry.first, (int)rz.first};
        copy_rr_(&xl, &yl, &zl, from.data.data(), &f[0], &f[1], &f[2], &f[3], &f[4], &f[5], data.data(), &t[0], &t[1], &t[2],
                 &t[3], &t[4], &t[5]);
    }
    // ri.rs:421-433 -> copy_mr_
    void copy_from_matr(Range rx, Range ry, size_t i_z, int copy_mod, const MatrixFull &from, Range fx, Range fy)
    {
        if (rlen(rx) != rlen(fx) || rlen(ry) != rlen(fy))
            throw std::runtime_error("Error: the data block for copy has different size between the matrix and ri 3D-tensor");
        int xl = (int)rlen(rx), yl = (int)rlen(ry), fxl = (int)from.size[0], fyl = (int)from.size[1], fxs = (int)fx.first,
            fys = (int)fy.first, tx = (int)size[0], ty = (int)size[1], tz = (int)size[2], txs = (int)rx.first,
            tys = (int)ry.first, t3 = (int)i_z;
        copy_mr_(&xl, &yl, from.data.data(), &fxl, &fyl, &fxs, &fys, data.data(), &tx, &ty, &tz, &txs, &tys, &t3, &copy_mod);
    }
    // d_P / J / K
    std::vector<double> ri_dp(const MatrixFull &dm) const
    {
        std::vector<double> d(size[2], 0.0);
        rb_check(rb_host_ri_dp(data.data(), dm.data.data(), d.data(), (int)size[0], (int)size[2]), "ri_dp");
        return d;
    }
    MatrixFull ri_j(const std::vector<double> &d) const
    {
        MatrixFull j = MatrixFull::make({size[0], size[0]}, 0.0);
        rb_check(rb_host_ri_j(data.data(), d.data(), j.data.data(), (int)size[0], (int)size[2]), "ri_j");
        return j;
    }
    MatrixFull ri_k(const MatrixFull &ct) const
    {
        MatrixFull k = MatrixFull::make({size[0], size[0]}, 0.0);
        rb_check(rb_host_ri_k(data.data(), ct.data.data(), (int)ct.size[1], k.data.data(), (int)size[0], (int)size[2]), "ri_k");
        return k;
    }
    // consumers of a P-fastest MO tensor (this = ri3mo[P, l, r]); ranges are half-open [lo, hi)
    MatrixFull ri_iajb(std::array<size_t, 2> la, std::array<size_t, 2> ra, std::array<size_t, 2> lb, std::array<size_t, 2> rb,
                       const RIFull *other = nullptr) const
    {
        const RIFull &b = other ? *other : *this;
        if (b.size[0] != size[0]) throw std::runtime_error("ri_iajb: the two MO tensors have different auxiliary dimensions");
        if (la[0] > la[1] || la[1] > size[1] || ra[0] > ra[1] || ra[1] > size[2] || lb[0] > lb[1] || lb[1] > b.size[1] ||
            rb[0] > rb[1] || rb[1] > b.size[2])
            throw std::runtime_error("ri_iajb: box outside the tensor");
        const size_t m = (la[1] - la[0]) * (ra[1] - ra[0]), n = (lb[1] - lb[0]) * (rb[1] - rb[0]);
        MatrixFull out = MatrixFull::make({m, n}, 0.0);
        rb_check(rb_host_ri_iajb((int)size[0], data.data(), (int)size[1], (int)size[2], (int)la[0], (int)(la[1] - la[0]),
                                 (int)ra[0], (int)(ra[1] - ra[0]), b.data.data(), (int)b.size[1], (int)b.size[2], (int)lb[0],
                                 (int)(lb[1] - lb[0]), (int)rb[0], (int)(rb[1] - rb[0]), out.data.data()), "ri_iajb");
        return out;
    }
    MatrixFull ri_mo_pq(std::array<size_t, 2> l, std::array<size_t, 2> r, const std::vector<double> *w = nullptr) const
    {
        if (l[0] > l[1] || l[1] > size[1] || r[0] > r[1] || r[1] > size[2]) throw std::runtime_error("ri_mo_pq: box outside the tensor");
        if (w && w->size() != (l[1] - l[0]) * (r[1] - r[0])) throw std::runtime_error("ri_mo_pq: one weight per MO pair of the box is needed");
        MatrixFull out = MatrixFull::make({size[0], size[0]}, 0.0);
        rb_check(rb_host_ri_mo_pq(data.data(), (int)size[0], (int)size[1], (int)size[2], (int)l[0], (int)(l[1] - l[0]), (int)r[0],
                                  (int)(r[1] - r[0]), w ? w->data() : nullptr, out.data.data()), "ri_mo_pq");
        return out;
    }

  private:
    RIFull tr(int which, std::array<size_t, 3> ns) const
    {
        RIFull out = make(ns, 0.0);
        if (!data.empty())
            rb_check(rb_host_ri_transpose(data.data(), (int64_t)size[0], (int64_t)size[1], (int64_t)size[2], which, out.data.data()), "transpose");
        return out;
    }
};

inline RIFull MatrixFull::to_rifull(size_t i, size_t j, size_t k) const
{
    if (i * j * k != data.size()) throw std::runtime_error("Error: the RIFull shape does not hold the matrix elements");
    RIFull r = RIFull::from_vec({i, j, k}, data);
    r.indicing = {1, i, j}; // the reference's quirk (matrixfull.rs:681-685): [1, i, j], not [1, i, i*j]
    return r;
}

// ERIFold4 (eri.rs:170-373): (ij|kl) with both pairs folded, column-major [npair, npair]
struct ERIFold4 {
    std::array<size_t, 2> size{0, 0};
    std::array<size_t, 2> indicing{0, 0};
    std::vector<double> data;
    static ERIFold4 make(std::array<size_t, 2> size, double v)
    {
        ERIFold4 t;
        t.size = size; t.indicing = {1, size[0]};
        t.data.assign(size[0] * size[1], v);
        return t;
    }
    // eri.rs:308-372 (mode 1) / 266-305 (mode 0)
    void chunk_copy_from_a_full_vector(const std::array<Range, 4> &r, const std::vector<double> &buf) { scatter(r, buf, 1); }
    void chunk_copy_from_local_erifull(size_t dim, Range d1, Range d2, Range d3, Range d4, const std::vector<double> &buf)
    {
        if (dim * (dim + 1) / 2 != size[0]) throw std::runtime_error("chunk_copy_from_local_erifull: dim does not match the tensor");
        scatter({d1, d2, d3, d4}, buf, 0);
    }

  private:
    void scatter(const std::array<Range, 4> &r, const std::vector<double> &buf, int mode)
    {
        if (buf.size() < rlen(r[0]) * rlen(r[1]) * rlen(r[2]) * rlen(r[3])) throw std::runtime_error("ERIFold4: the local block is too short");
        rb_check(rb_host_erifold4_chunk_copy(data.data(), (int64_t)size[0], (int64_t)size[1], (int)r[0].first, (int)rlen(r[0]),
                                             (int)r[1].first, (int)rlen(r[1]), (int)r[2].first, (int)rlen(r[2]), (int)r[3].first,
                                             (int)rlen(r[3]), buf.data(), mode), "ERIFold4 chunk copy");
    }
};

// matrix_blas_lapack.rs:256-278: allocates C with the op-dependent shape
inline MatrixFull _dgemm_full_new(const MatrixFull &a, char opa, const MatrixFull &b, char opb, double alpha, double beta);
// matrix_blas_lapack.rs:354-378: no checks, like the reference
inline void _dsymm(const MatrixFull &a, const MatrixFull &b, MatrixFull &c, char side, char uplo, double alpha, double beta)
{
    const size_t m = c.size[0], n = c.size[1];
    const size_t lda = (side == 'L' || side == 'l') ? m : n;
    rb_check(rb_host_dsymm(side, uplo, (int)m, (int)n, alpha, a.data.data(), (int)std::max<size_t>(lda, 1), b.data.data(),
                           (int)std::max<size_t>(m, 1), beta, c.data.data(), (int)std::max<size_t>(m, 1)), "_dsymm");
}

// matrix_blas_lapack.rs:180-252
inline void _dgemm_full(const MatrixFull &a, char opa, const MatrixFull &b, char opb, MatrixFull &c, double alpha, double beta)
{
    size_t m = opa == 'N' ? a.size[0] : a.size[1], k = opa == 'N' ? a.size[1] : a.size[0];
    size_t n = opb == 'N' ? b.size[1] : b.size[0], kb = opb == 'N' ? b.size[0] : b.size[1];
    if (!((opa == 'N' || opa == 'T') && (opb == 'N' || opb == 'T')) || k != kb || m != c.size[0] || n != c.size[1])
        throw std::runtime_error("ERROR:: _dgemm_full shape mismatch");
    size_t lda = opa == 'N' ? std::max<size_t>(m, 1) : std::max<size_t>(k, 1);
    size_t ldb = opb == 'N' ? std::max<size_t>(k, 1) : std::max<size_t>(n, 1);
    rb_check(rb_host_dgemm(opa, opb, (int)m, (int)n, (int)k, alpha, a.data.data(), (int)lda, b.data.data(), (int)ldb, beta,
                           c.data.data(), (int)std::max<size_t>(m, 1)), "_dgemm_full");
}
inline MatrixFull _dgemm_full_new(const MatrixFull &a, char opa, const MatrixFull &b, char opb, double alpha, double beta)
{
    const size_t m = opa == 'N' ? a.size[0] : a.size[1], n = opb == 'N' ? b.size[1] : b.size[0];
    MatrixFull c = MatrixFull::make({m, n}, 0.0);
    _dgemm_full(a, opa, b, opb, c, alpha, beta);
    return c;
}
// matrix_blas_lapack.rs:392-413
inline void _dsyrk(const MatrixFull &a, MatrixFull &c, char uplo, char trans, double alpha, double beta)
{
    if (c.size[0] != c.size[1]) throw std::runtime_error("matr_b should be symmetric");
    bool is_n = trans == 'N' || trans == 'n';
    size_t n = c.size[0], k = is_n ? a.size[1] : a.size[0];
    size_t lda = is_n ? std::max<size_t>(n, 1) : std::max<size_t>(k, 1);
    rb_check(rb_host_dsyrk(uplo, trans, (int)n, (int)k, alpha, a.data.data(), (int)lda, beta, c.data.data(), (int)std::max<size_t>(n, 1)), "_dsyrk");
}
// matrix_blas_lapack.rs:38-70
inline void _dgemv(const MatrixFull &a, const std::vector<double> &x, std::vector<double> &y, char trans, double alpha,
                   double beta, int incx, int incy)
{
    size_t m = a.size[0], n = a.size[1];
    bool is_n = trans == 'N' || trans == 'n';
    size_t lx = is_n ? n : m, ly = is_n ? m : n;
    if (x.size() != 1 + (lx - 1) * (size_t)std::abs(incx) || y.size() != 1 + (ly - 1) * (size_t)std::abs(incy))
        throw std::runtime_error("ERROR:: _dgemv length mismatch");
    rb_check(rb_host_dgemv(trans, (int)m, (int)n, alpha, a.data.data(), (int)std::max<size_t>(m, 1), x.data(), incx, beta, y.data(), incy), "_dgemv");
}

} // namespace rest_tensors
