// C++ host-mirror test (compiled and run by tests/test_gpu_cpp_host.py on a GPU box): the reference's own worked
// examples through rest_tensors.hpp -> C ABI -> CUDA, checked against the golden vectors and the CPU oracle.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../rest_tensors_b200/host/rest_tensors.hpp"

extern "C" { // oracle (test infrastructure)
void orc_fill_linear(double *v, int64_t n, uint64_t seed, uint64_t idx0, double scale);
void orc_ri_ao2mo_f(const double *c, const double *ri, double *mo, int ns, int nb, int nx);
void orc_ri_dp(const double *ri, const double *dm, double *d, int nb, int nx);
void orc_ri_j(const double *ri, const double *d, double *j, int nb, int nx);
void orc_ri_k(const double *ri, const double *ct, double *k, int nb, int no, int nx);
int orc_dsyev(char jobz, int n, double *a, double *w);
int orc_load_blas(const char *path);
void orc_ri_iajb(int np, const double *mo_a, int nl_a, int l0a, int lla, int r0a, int rla, const double *mo_b, int nl_b,
                 int l0b, int llb, int r0b, int rlb, double *out);
void orc_ri_mo_pq(const double *mo_a, int npa, const double *mo_b, int npb, int nl, int l0, int ll, int r0, int rl,
                  const double *w, double *out);
void orc_ri_transpose(const double *in, int64_t I, int64_t J, int64_t K, int which, double *out);
}
using namespace rest_tensors;

static int fails = 0;
#define CHECK(cond, what) do { if (!(cond)) { std::printf("FAIL: %s\n", what); ++fails; } } while (0)

static double rel_err(const std::vector<double> &x, const std::vector<double> &y)
{
    double num = 0, den = 0;
    for (size_t i = 0; i < y.size(); ++i) { num = std::fmax(num, std::fabs(x[i] - y[i])); den = std::fmax(den, std::fabs(y[i])); }
    return den > 0 ? num / den : num;
}
static std::vector<double> fill(size_t n, uint64_t seed, double scale = 1.0)
{
    std::vector<double> v(n);
    orc_fill_linear(v.data(), (int64_t)n, seed, 0, scale);
    return v;
}

int main()
{
    // GV4: iter_matrixupper of 4x4 (1..16)  (reference src/matrix/mod.rs:437-452)
    std::vector<double> a16(16);
    for (int i = 0; i < 16; ++i) a16[i] = i + 1;
    auto up = MatrixFull::from_vec({4, 4}, a16).to_matrixupper();
    CHECK((up.data == std::vector<double>{1, 5, 6, 9, 10, 11, 13, 14, 15, 16}), "GV4 pack order");
    // GV5: MatrixUpper -> MatrixFull  (reference matrix_blas_lapack.rs:507-513)
    auto full = MatrixUpper::from_vec(6, {4, 12, 37, -16, -43, 98}).to_matrixfull();
    CHECK(full.has_value() && (full->data == std::vector<double>{4, 12, -16, 12, 37, -43, -16, -43, 98}), "GV5 unpack");
    CHECK(!MatrixUpper::from_vec(4, {1, 2, 3, 4}).to_matrixfull().has_value(), "non-triangular length -> None");
    // GV7: transpose of 3x4 (1..12)
    std::vector<double> a12(12);
    for (int i = 0; i < 12; ++i) a12[i] = i + 1;
    auto t = MatrixFull::from_vec({3, 4}, a12).transpose();
    CHECK(t.size[0] == 4 && t.data[8] == 3 && t.data[9] == 6 && t.data[10] == 9 && t.data[11] == 12, "GV7 transpose");
    // GV9: benches/bench_tensors.rs -- every element 200
    auto mo = RIFull::make({10, 10, 20}, 2.0).ao2mo_v02(MatrixFull::make({10, 10}, 1.0));
    bool all200 = mo.size[0] == 20;
    for (double v : mo.data) all200 = all200 && v == 200.0;
    CHECK(all200, "GV9 ao2mo bench case");
    // GV10: the four transposes of RIFull[3,2,2] 0..12 vs oracle
    std::vector<double> r12(12);
    for (int i = 0; i < 12; ++i) r12[i] = i;
    auto ri = RIFull::from_vec({3, 2, 2}, r12);
    RIFull trs[4] = {ri.transpose_jik(), ri.transpose_jki(), ri.transpose_kji(), ri.transpose_ikj()};
    for (int w = 0; w < 4; ++w) {
        std::vector<double> ref(12);
        orc_ri_transpose(r12.data(), 3, 2, 2, w, ref.data());
        CHECK(trs[w].data == ref, "GV10 RIFull transposes");
    }
    // ao2mo / d_P / J / K vs the oracle on a random case
    const int nb = 48, nx = 30, no = 7;
    auto riv = fill((size_t)nb * nb * nx, 2), c = fill((size_t)nb * nb, 3, 1.0 / std::sqrt((double)nb));
    auto R = RIFull::from_vec({(size_t)nb, (size_t)nb, (size_t)nx}, riv);
    auto C = MatrixFull::from_vec({(size_t)nb, (size_t)nb}, c);
    std::vector<double> ref((size_t)nx * nb * nb);
    orc_ri_ao2mo_f(c.data(), riv.data(), ref.data(), nb, nb, nx);
    CHECK(rel_err(R.ao2mo(C).data, ref) < 1e-10, "ao2mo vs oracle");
    auto dm = fill((size_t)nb * nb, 5);
    std::vector<double> dref(nx), jref((size_t)nb * nb), kref((size_t)nb * nb);
    orc_ri_dp(riv.data(), dm.data(), dref.data(), nb, nx);
    orc_ri_j(riv.data(), dref.data(), jref.data(), nb, nx);
    std::vector<double> ct(c.begin(), c.begin() + (size_t)nb * no);
    orc_ri_k(riv.data(), ct.data(), kref.data(), nb, no, nx);
    CHECK(rel_err(R.ri_dp(MatrixFull::from_vec({(size_t)nb, (size_t)nb}, dm)), dref) < 1e-10, "d_P vs oracle");
    CHECK(rel_err(R.ri_j(dref).data, jref) < 1e-10, "J vs oracle");
    CHECK(rel_err(R.ri_k(MatrixFull::from_vec({(size_t)nb, (size_t)no}, ct)).data, kref) < 1e-10, "K vs oracle");
    // consumers of ri3mo: (ia|jb) block and the RPA-type auxiliary-basis matrix
    {
        const int np = 40, nl = 5, nr = 7;
        auto mov = fill((size_t)np * nl * nr, 8), wv = fill(3 * 4, 9);
        auto M = RIFull::from_vec({(size_t)np, (size_t)nl, (size_t)nr}, mov);
        std::vector<double> gref(12 * 35), pref((size_t)np * np);
        orc_ri_iajb(np, mov.data(), nl, 1, 3, 2, 4, mov.data(), nl, 0, 5, 0, 7, gref.data());
        orc_ri_mo_pq(mov.data(), np, mov.data(), np, nl, 1, 3, 2, 4, wv.data(), pref.data());
        CHECK(rel_err(M.ri_iajb({1, 4}, {2, 6}, {0, 5}, {0, 7}).data, gref) < 1e-10, "(ia|jb) vs oracle");
        CHECK(rel_err(M.ri_mo_pq({1, 4}, {2, 6}, &wv).data, pref) < 1e-10, "mo_pq vs oracle");
        bool thr = false;
        try { M.ri_iajb({0, 6}, {0, 7}, {0, 5}, {0, 7}); } catch (const std::runtime_error &) { thr = true; }
        CHECK(thr, "ri_iajb box outside the tensor must throw");
    }
    // eigen-solver behind lapack_dsyev: eigenvalues vs residual / orthogonality invariants (and vs LAPACK when OPENBLAS_PATH is set)
    {
        const int n = 50;
        auto av = fill((size_t)n * n, 10);
        for (int j = 0; j < n; ++j) for (int i = 0; i < j; ++i) av[(size_t)i + (size_t)j * n] = av[(size_t)j + (size_t)i * n]; // symmetric
        auto A = MatrixFull::from_vec({(size_t)n, (size_t)n}, av);
        auto zw = A.lapack_dsyev();
        double res = 0.0, orth = 0.0, wmax = 0.0;
        for (int c = 0; c < n; ++c) wmax = std::max(wmax, std::fabs(zw.second[c]));
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < n; ++r) {
                double s = 0.0;
                for (int k = 0; k < n; ++k) s += av[(size_t)r + (size_t)k * n] * zw.first.data[(size_t)k + (size_t)c * n];
                res = std::max(res, std::fabs(s - zw.second[c] * zw.first.data[(size_t)r + (size_t)c * n]));
            }
        for (int c = 0; c < n; ++c)
            for (int d = 0; d < n; ++d) {
                double s = 0.0;
                for (int k = 0; k < n; ++k) s += zw.first.data[(size_t)k + (size_t)c * n] * zw.first.data[(size_t)k + (size_t)d * n];
                orth = std::max(orth, std::fabs(s - (c == d ? 1.0 : 0.0)));
            }
        CHECK(res < 1e-11 * wmax * n && orth < 1e-12 * n, "lapack_dsyev residual / orthogonality");
        if (const char *blas = std::getenv("OPENBLAS_PATH")) {
            if (orc_load_blas(blas) == 0) {
                std::vector<double> aref = av, wref(n);
                CHECK(orc_dsyev('V', n, aref.data(), wref.data()) == 0, "oracle dsyev");
                CHECK(rel_err(zw.second, wref) < 1e-10, "lapack_dsyev eigenvalues vs LAPACK");
            }
        }
        auto S = MatrixFull::from_vec({(size_t)n, (size_t)n}, av);
        for (int i = 0; i < n; ++i) S.data[(size_t)i * (n + 1)] += 2.0 * n;   // diagonally dominant -> positive definite
        auto X = S.lapack_power(-0.5, 1e-10);
        double dev = 0.0;   // X S X = I
        std::vector<double> t((size_t)n * n, 0.0), u((size_t)n * n, 0.0);
        for (int j = 0; j < n; ++j) for (int k = 0; k < n; ++k) for (int i = 0; i < n; ++i) t[(size_t)i + (size_t)j * n] += X.data[(size_t)i + (size_t)k * n] * S.data[(size_t)k + (size_t)j * n];
        for (int j = 0; j < n; ++j) for (int k = 0; k < n; ++k) for (int i = 0; i < n; ++i) u[(size_t)i + (size_t)j * n] += t[(size_t)i + (size_t)k * n] * X.data[(size_t)k + (size_t)j * n];
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) dev = std::max(dev, std::fabs(u[(size_t)i + (size_t)j * n] - (i == j ? 1.0 : 0.0)));
        CHECK(dev < 1e-11 * n, "lapack_power(-0.5): X S X = I");
    }
    // round-2 additions to the mirror: occ-vir ao2mo == sub-block of the square one, _dgemm_full_new / ddot / lapack_dgemm, views,
    // metadata reshapes, ERIFold4 shell-quartet scatter
    {
        const int nb = 24, nx = 10, no = 5;
        auto ri = RIFull::from_vec({(size_t)nb, (size_t)nb, (size_t)nx}, fill((size_t)nb * nb * nx, 31));
        auto c = MatrixFull::from_vec({(size_t)nb, (size_t)nb}, fill((size_t)nb * nb, 32, 0.2));
        auto sq = ri.ao2mo(c);
        auto cl = MatrixFull::from_vec({(size_t)nb, (size_t)no}, std::vector<double>(c.data.begin(), c.data.begin() + nb * no));
        auto cr = MatrixFull::from_vec({(size_t)nb, (size_t)(nb - no)}, std::vector<double>(c.data.begin() + nb * no, c.data.end()));
        auto ov = ri.ao2mo_rect(cl, cr);
        bool same = ov.size[0] == (size_t)nx && ov.size[1] == (size_t)no && ov.size[2] == (size_t)(nb - no);
        for (int b = 0; b < nb - no && same; ++b)
            for (int a = 0; a < no && same; ++a)
                for (int p = 0; p < nx; ++p)
                    same = same && std::fabs(ov.data[p + (size_t)nx * (a + (size_t)no * b)] - sq.data[p + (size_t)nx * (a + (size_t)nb * (b + no))]) <= 1e-12;
        CHECK(same, "ao2mo_rect == occ-vir block of the square transform");
        {   // streaming pass: full, upper pairs, upper pairs from half-uploaded symmetric slabs
            std::vector<double> sym(ri.data);
            for (int p = 0; p < nx; ++p) for (int j = 0; j < nb; ++j) for (int i = j + 1; i < nb; ++i)
                sym[i + (size_t)nb * (j + (size_t)nb * p)] = sym[j + (size_t)nb * (i + (size_t)nb * p)];
            auto rs = RIFull::from_vec({(size_t)nb, (size_t)nb, (size_t)nx}, sym);
            auto dmat = MatrixFull::from_vec({(size_t)nb, (size_t)nb}, fill((size_t)nb * nb, 34, 0.1));
            auto full = rs.ao2mo_jk(c, dmat, cl), up = rs.ao2mo_jk(c, dmat, cl, true), sy = rs.ao2mo_jk(c, dmat, cl, true, true);
            bool okp = up.ri3mo.size() == (size_t)nx * nb * (nb + 1) / 2;
            for (int b = 0; b < nb && okp; ++b) for (int a = 0; a <= b && okp; ++a) for (int p = 0; p < nx; ++p)
                okp = okp && up.ri3mo[p + (size_t)nx * ((size_t)b * (b + 1) / 2 + a)] == full.ri3mo[p + (size_t)nx * (a + (size_t)nb * b)];
            CHECK(okp, "ao2mo_jk(upper) == a <= b pairs of the full pass");
            CHECK(sy.ri3mo == up.ri3mo && sy.k.data == up.k.data && sy.j.data == up.j.data && sy.d == up.d,
                  "ao2mo_jk(upper, symmetric_slabs) == ao2mo_jk(upper) bit for bit");
            CHECK(rel_err(full.ri3mo, rs.ao2mo(c).data) < 1e-13, "ao2mo_jk ri3mo == ao2mo");
        }
        auto prod = _dgemm_full_new(c, 'T', c, 'N', 1.0, 0.0);
        auto dd = c.transpose().ddot(c);
        CHECK(dd.has_value() && rel_err(dd->data, prod.data) < 1e-13, "ddot == _dgemm_full_new");
        CHECK(!c.ddot(MatrixFull::make({3, 3}, 1.0)).has_value(), "ddot shape mismatch -> None");
        auto ijk = ri.rifull_to_matfull_ij_k();
        CHECK(ijk.size[0] == (size_t)nb * nb && ijk.size[1] == (size_t)nx && ijk.data == ri.data, "rifull_to_matfull_ij_k");
        auto back = ijk.to_rifull(nb, nb, nx);
        CHECK(back.indicing[2] == (size_t)nb && back.data == ri.data, "to_rifull keeps the reference's indicing quirk");
        auto runs = ri.get_slices({2, 7}, {1, 3}, {4, 6});
        CHECK(runs.size() == 4 && runs[0].second == 5 && runs[1].first == ri.data.data() + 2 + 2 * nb + 4 * nb * nb, "get_slices run order");
        const size_t dim = 7, npair = dim * (dim + 1) / 2;
        auto g = fill(dim * dim * dim * dim, 33);
        auto eri = ERIFold4::make({npair, npair}, -1.0);
        std::array<Range, 4> whole = {Range{0, dim}, Range{0, dim}, Range{0, dim}, Range{0, dim}};
        eri.chunk_copy_from_a_full_vector(whole, g);
        bool ok4 = true;
        for (size_t l = 0; l < dim; ++l) for (size_t k = 0; k <= l; ++k) for (size_t j = 0; j < dim; ++j) for (size_t i = 0; i <= j; ++i)
            ok4 = ok4 && eri.data[(j * (j + 1) / 2 + i) + npair * (l * (l + 1) / 2 + k)] == g[i + dim * (j + dim * (k + dim * l))];
        CHECK(ok4, "ERIFold4 chunk_copy_from_a_full_vector (whole tensor as one block)");
    }
    // panics
    bool threw = false;
    try { RIFull::from_vec({3, 2, 2}, std::vector<double>(11)); } catch (const std::runtime_error &) { threw = true; }
    CHECK(threw, "from_vec too short must throw");
    threw = false;
    try { MatrixFull::make({2, 3}, 0.0).to_matrixupper(); } catch (const std::runtime_error &) { threw = true; }
    CHECK(threw, "to_matrixupper of a non-square matrix must throw");
    if (fails == 0) std::printf("CPP_HOST_MIRROR_OK\n");
    return fails == 0 ? 0 : 1;
}
