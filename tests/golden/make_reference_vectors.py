"""Writes tests/golden/reference_vectors.json.

The reference (Rust + Fortran + un-vendored OpenBLAS) cannot be built or imported in this image, so these
vectors are TRANSCRIBED from the `assert_eq!` expectations of the reference's own doc-tests (GV1-GV8), with the
reference file:line of each; GV9/GV10 are derived by hand from the reference's bench / print-only test inputs
(the reference asserts nothing for them) and are flagged "derived".  Run from the repo root:
    python tests/golden/make_reference_vectors.py
"""
import json
import os

V = {
    "GV1": {
        "source": "src/matrix/matrix_blas_lapack.rs:78-98 and src/external_libs/mod.rs:34-58 (assert_eq)",
        "pinned_by_reference": True,
        "a": [float(x) for x in range(1, 10)], "size_a": [3, 3],
        "b": [float(x) for x in range(6, 15)], "size_b": [3, 3],
        "c_fill": 2.0, "size_c": [3, 3],
        "sub_a": [[1, 3], [1, 3]], "opa": "N", "sub_b": [[0, 2], [0, 2]], "opb": "N",
        "sub_c": [[1, 3], [0, 2]], "alpha": 1.0, "beta": 1.0,
        "expect_block": [88.0, 101.0, 127.0, 146.0],
    },
    "GV2": {
        "source": "src/matrix/matrix_blas_lapack.rs:101-109 (assert_eq)",
        "pinned_by_reference": True,
        "a": [float(x) for x in range(1, 10)], "size_a": [3, 3],
        "b": [float(x) for x in range(6, 15)], "size_b": [3, 3],
        "c_fill": 2.0, "size_c": [3, 3],
        "sub_a": [[1, 3], [1, 2]], "opa": "T", "sub_b": [[0, 2], [0, 2]], "opb": "N",
        "sub_c": [[1, 2], [0, 2]], "alpha": 1.0, "beta": 1.0,
        "expect_block": [74.0, 107.0],
    },
    "GV3": {
        "source": "src/matrix/matrix_blas_lapack.rs:111-119 (assert_eq)",
        "pinned_by_reference": True,
        "a": [float(x) for x in range(1, 10)], "size_a": [3, 3],
        "b": [float(x) for x in range(6, 15)], "size_b": [3, 3],
        "c_fill": 2.0, "size_c": [3, 3],
        "sub_a": [[1, 3], [0, 2]], "opa": "N", "sub_b": [[0, 1], [0, 2]], "opb": "T",
        "sub_c": [[0, 2], [0, 1]], "alpha": 1.0, "beta": 1.0,
        "expect_block": [59.0, 74.0],
    },
    "GV4": {
        "source": "src/matrix/mod.rs:437-452 iter_matrixupper of 4x4 (1..16) (assert_eq)",
        "pinned_by_reference": True,
        "full": [float(x) for x in range(1, 17)], "n": 4,
        "packed": [1.0, 5.0, 6.0, 9.0, 10.0, 11.0, 13.0, 14.0, 15.0, 16.0],
    },
    "GV5": {
        "source": "src/matrix/matrix_blas_lapack.rs:290-295 and 507-513 (to_matrixfull() results drawn in the doc-test "
                  "comments; the doc-tests then assert _dsyev/_dpotrf results computed FROM these matrices)",
        "pinned_by_reference": True,
        "cases": [
            {"packed": [1.0, 2.0, 3.0, 4.0, 5.0, 6.0],
             "full_colmajor": [1.0, 2.0, 4.0, 2.0, 3.0, 5.0, 4.0, 5.0, 6.0]},
            {"packed": [4.0, 12.0, 37.0, -16.0, -43.0, 98.0],
             "full_colmajor": [4.0, 12.0, -16.0, 12.0, 37.0, -43.0, -16.0, -43.0, 98.0]},
        ],
    },
    "GV7": {
        "source": "src/matrix/matrixfull.rs:531-551 transpose of 3x4 (1..12): row 2 <-> column 2 = [3,6,9,12]",
        "pinned_by_reference": True,
        "data": [float(x) for x in range(1, 13)], "size": [3, 4],
        "transposed_column_2": [3.0, 6.0, 9.0, 12.0],
    },
    "GV8": {
        "source": "src/matrix/matrix_blas_lapack.rs:285-317 _dsyev doc-test: MatrixUpper 1..6 -> full 3x3; eigenvectors and "
                  "eigenvalues asserted with sum of squared differences < 10E-7 (jobz 'V' and 'N')",
        "pinned_by_reference": True,
        "packed": [float(x) for x in range(1, 7)], "n": 3,
        "eigenvectors": [-0.6827362941552275, -0.38559063640162244, 0.6206375864887483,
                         0.6202872696512856, -0.7547821848190948, 0.21341873532627673,
                         -0.3861539275363389, -0.5306823104260265, -0.7544941548144386],
        "eigenvalues": [-1.5066326307865059, -0.05739624271478554, 11.564028873501286],
        "tolerance_sum_sq": 1e-6,
    },
    "GV9": {
        "source": "benches/bench_tensors.rs:5-9: RIFull::new([10,10,20],2.0).ao2mo_v02(MatrixFull::new([10,10],1.0)); "
                  "every output = 2*10*10 (derived; the bench asserts nothing)",
        "pinned_by_reference": False,
        "ri_size": [10, 10, 20], "ri_fill": 2.0, "c_size": [10, 10], "c_fill": 1.0, "expect_every": 200.0,
    },
    "GV10": {
        "source": "src/ri.rs:437-448 print-only test: RIFull [3,2,2] data 0..12 (derived from the code paths ri.rs:227-294)",
        "pinned_by_reference": False,
        "size": [3, 2, 2], "data": [float(x) for x in range(12)],
        "jik": [0, 3, 1, 4, 2, 5, 6, 9, 7, 10, 8, 11],
        "jki": [0, 3, 6, 9, 1, 4, 7, 10, 2, 5, 8, 11],
        "kji": [0, 6, 3, 9, 1, 7, 4, 10, 2, 8, 5, 11],
        "ikj": [0, 1, 2, 6, 7, 8, 3, 4, 5, 9, 10, 11],
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(V, f, indent=1)
    print("wrote", out)
