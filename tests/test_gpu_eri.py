"""ERIFold4 (reference src/eri.rs:170-373) -- the chunk scatters that fold libcint shell-quartet blocks into
the packed [npair, npair] tensor: CUDA (device and host-pointer forms through the C ABI) against the CPU restatement of the
reference's loops, bit for bit, plus a whole-tensor assembly from shell blocks against the direct definition."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ranges_cases(dim):
    shells = [(0, 3), (3, 4), (4, 9), (9, dim)]
    cases = []
    for a, b, c, d in itertools.product(range(len(shells)), repeat=4):
        cases.append((shells[a], shells[b], shells[c], shells[d]))
    return shells, cases


def test_chunk_copies_vs_reference_loops(rt, oracle):
    dim = 12
    npair = dim * (dim + 1) // 2
    shells, cases = _ranges_cases(dim)
    for mode in (0, 1):
        for r in cases:
            need = int(np.prod([b - a for a, b in r]))
            buf = oracle.fill_linear(need, 91)
            ref = oracle.fill_linear(npair * npair, 92)
            t = rt.ERIFold4.from_vec([npair, npair], ref.copy())
            if mode == 0:
                ok = oracle.erifold4_chunk_copy_local(ref, [npair, npair], dim, r, buf)
            else:
                ok = oracle.erifold4_chunk_copy_full(ref, [npair, npair], r, buf)
            if not ok:      # the reference panics (a slice leaves the tensor): the C ABI must refuse as well
                with pytest.raises(rt.RestB200Error):
                    (t.chunk_copy_from_local_erifull(dim, *r, buf) if mode == 0 else t.chunk_copy_from_a_full_vector(list(r), buf))
                continue
            if mode == 0:
                t.chunk_copy_from_local_erifull(dim, *r, buf)
            else:
                t.chunk_copy_from_a_full_vector(list(r), buf)
            assert np.array_equal(t.data, ref), (mode, r)


def test_whole_tensor_assembled_from_shell_quartets(rt, oracle):
    """Folding every shell quartet with i-shell <= j-shell and k-shell <= l-shell reproduces eri[(i,j),(k,l)] = g(i,j,k,l) for all
    i <= j, k <= l -- the way REST fills the tensor from libcint (chunk_copy_from_a_full_vector per quartet)."""
    dim = 11
    npair = dim * (dim + 1) // 2
    bounds = [0, 2, 5, 6, dim]
    shells = list(zip(bounds[:-1], bounds[1:]))
    g = oracle.fill_linear(dim ** 4, 93).reshape((dim,) * 4, order="F")     # g[i, j, k, l]
    t = rt.ERIFold4.new([npair, npair], float("nan"))
    for (sa, sb, sc, sd) in itertools.product(range(len(shells)), repeat=4):
        if sa > sb or sc > sd:
            continue
        r = [shells[sa], shells[sb], shells[sc], shells[sd]]
        blk = g[r[0][0]:r[0][1], r[1][0]:r[1][1], r[2][0]:r[2][1], r[3][0]:r[3][1]]
        t.chunk_copy_from_a_full_vector(r, np.ascontiguousarray(blk.reshape(-1, order="F")))
    m = t.data.reshape((npair, npair), order="F")
    for j in range(dim):
        for i in range(j + 1):
            for l in range(dim):
                for k in range(l + 1):
                    assert m[j * (j + 1) // 2 + i, l * (l + 1) // 2 + k] == g[i, j, k, l]
    assert not np.isnan(t.data).any()
    col = t.get_reducing_matrix(4)
    assert col.size == npair and np.shares_memory(col.data, t.data)


def test_device_form_large_block(ctx, oracle):
    """device-resident tensor, one big block (a whole 40-function chunk): compare with an independent torch construction"""
    dim = 40
    npair = dim * (dim + 1) // 2
    r = ((0, 24), (8, 40), (3, 30), (10, 40))
    li, lj, lk, ll = [b - a for a, b in r]
    buf = ctx.empty(li * lj * lk * ll); ctx.fill_linear(buf, buf.numel(), 94, 0, 1.0)
    eri = ctx.empty(npair * npair); ctx.fill_linear(eri, eri.numel(), 95, 0, 1.0)
    want = eri.clone().view(npair, npair)                      # [col (kl)][row (ij)]
    b4 = buf.view(ll, lk, lj, li)
    i = torch.arange(r[0][0], r[0][1], device=buf.device); j = torch.arange(r[1][0], r[1][1], device=buf.device)
    k = torch.arange(r[2][0], r[2][1], device=buf.device); l = torch.arange(r[3][0], r[3][1], device=buf.device)
    L, K, J, I = torch.meshgrid(l, k, j, i, indexing="ij")
    keep = (K <= L) & (I <= J)
    want[(L * (L + 1) // 2 + K)[keep], (J * (J + 1) // 2 + I)[keep]] = b4[keep]
    ctx.erifold4_chunk_copy(eri, npair, npair, npair, r, buf, 0)
    assert torch.equal(eri.view(npair, npair), want)
    with pytest.raises(Exception):
        ctx.erifold4_chunk_copy(eri, npair, npair, npair, ((0, 24), (8, 41), (3, 30), (10, 40)), buf, 0)   # j = 40 is outside
