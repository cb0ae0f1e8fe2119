"""CUDA-graph recording of a repeated call sequence (rb_graph_begin / rb_graph_end / rb_graph_launch): the per-iteration work of
the reference's SCF loop -- d_P, J, K (/root/reference/src/ri.rs) and, for small systems, ao2mo --
recorded once and replayed with one launch.  A replay must give the bits the calls give when issued one by one, follow the CONTENTS
of its input buffers, survive later calls that grow the workspaces, and refuse what cannot be recorded, loudly."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Step:
    """ao2mo + d_P + J + K on a resident ri3ao shard with every buffer allocated up front"""

    def __init__(self, ctx, nb, nx, no, seed=3):
        from rest_tensors_b200.device import ShardedRI
        self.ctx, self.nb, self.nx, self.no = ctx, nb, nx, no
        self.sh = ShardedRI(ctx, nb, nx).fill_synthetic()
        n2 = nb * nb
        self.c = ctx.empty(n2); self.ct = ctx.empty(nb * no); self.dm = ctx.empty(n2)
        self.mo = ctx.empty(nx * n2); self.d = ctx.empty(nx); self.j = ctx.empty(n2); self.k = ctx.empty(n2)
        self.set_inputs(seed)

    def set_inputs(self, seed):
        ctx, nb, no = self.ctx, self.nb, self.no
        ctx.fill_linear(self.c, nb * nb, seed, 0, nb ** -0.5)
        ctx.fill_linear(self.ct, nb * no, seed + 1, 0, nb ** -0.5)
        ctx.fill_linear(self.dm, nb * nb, seed + 2, 0, 1.0 / nb)

    def run(self):
        sh, nb, no = self.sh, self.nb, self.no
        sh.ao2mo(self.c, nb, self.c, nb, out=self.mo)
        sh.dp(self.dm, out=self.d)
        sh.j(self.d, out=self.j)
        sh.k(self.ct, no, out=self.k)

    def outputs(self):
        return [t.clone() for t in (self.mo, self.d, self.j, self.k)]

    def clear(self):
        for t in (self.mo, self.d, self.j, self.k):
            t.fill_(float("nan"))


@pytest.fixture()
def side():
    """a fresh context on a non-default stream (the legacy default stream cannot be captured)"""
    from rest_tensors_b200.device import Context
    torch.cuda.set_device(0)              # earlier multi-GPU tests may have left another device current
    s = torch.cuda.Stream(device=0)
    with torch.cuda.stream(s):
        ctx = Context(0)
        yield ctx
        torch.cuda.synchronize()
        ctx.close()


@pytest.mark.parametrize("nb,nx,no", [(100, 400, 20), (264, 90, 21), (45, 77, 7)])
def test_replay_matches_the_calls_issued_one_by_one(side, nb, nx, no):
    ctx = side
    st = Step(ctx, nb, nx, no)
    st.run()                                  # sizes the workspaces
    want = st.outputs()
    with ctx.record() as rec:
        st.run()
    g = rec.graph
    assert g.kernels >= 4
    for _ in range(3):
        st.clear()
        n0 = ctx.launches
        g.launch()
        assert ctx.launches - n0 == g.kernels
        for a, b in zip(st.outputs(), want):
            assert torch.equal(a, b)
    # the recording follows the contents of its inputs
    st.set_inputs(11)
    st.run(); want2 = st.outputs()
    assert not torch.equal(want2[3], want[3])
    st.clear(); g.launch()
    for a, b in zip(st.outputs(), want2):
        assert torch.equal(a, b)
    g.close()


def test_recording_survives_workspace_growth_and_other_calls(side):
    c2 = side
    st = Step(c2, 100, 64, 20)
    st.run(); want = st.outputs()
    with c2.record() as rec:
        st.run()
    # a much larger split-K product on the same context: the partial workspace grows, the recording keeps its own block
    m, k = 600, 40000
    a = c2.empty(m * k); cbig = c2.empty(m * m)
    c2.fill_linear(a, m * k, 5, 0, 1.0)
    c2.dsyrk("U", "N", m, k, 1.0, a, m, 0.0, cbig, m)
    c2.poison_workspaces()
    st.clear(); rec.graph.launch()
    for x, y in zip(st.outputs(), want):
        assert torch.equal(x, y)
    rec.graph.close()


def test_what_cannot_be_recorded_is_refused_loudly(side):
    from rest_tensors_b200._lib import RestB200Error, lib
    ctx = side
    st = Step(ctx, 45, 30, 7)
    st.run(); want = st.outputs()
    n = 12
    s = ctx.empty(n * n); ctx.fill_linear(s, n * n, 2, 0, 1.0)
    w = ctx.empty(n); z = ctx.empty(n * n)
    with ctx.record() as rec:
        st.run()
        with pytest.raises(RestB200Error, match="cannot be recorded"):
            ctx.dsyev("V", "U", n, s, n, w, z, n)
        with pytest.raises(RestB200Error, match="cannot be recorded"):
            ctx.sync()
    st.clear(); rec.graph.launch()               # the refused calls left the recording intact
    for x, y in zip(st.outputs(), want):
        assert torch.equal(x, y)
    rec.graph.close()
    # a workspace that would have to grow while recording
    big = Step(ctx, 128, 300, 16)
    with pytest.raises(RestB200Error, match="before rb_graph_begin"):
        with ctx.record():
            big.run()
    big.run()                                    # the context is usable afterwards
    assert bool(torch.isfinite(big.k).all())
    # no recording open
    h = C.c_void_p()
    assert lib.rb_graph_end(ctx.h, C.byref(h)) != 0


def test_default_stream_is_refused(ctx):
    from rest_tensors_b200._lib import RestB200Error
    torch.cuda.set_device(0)
    ctx.bind_stream()
    if torch.cuda.current_stream().cuda_stream != 0:
        pytest.skip("torch's current stream is not the default stream here")
    with pytest.raises(RestB200Error, match="default stream"):
        with ctx.record():
            pass


def test_host_copies_can_be_part_of_the_recording(side):
    """upload of D and C~ from pinned host memory, d_P + J + K, download of J and K: one launch per SCF iteration"""
    from rest_tensors_b200._lib import check, lib
    ctx = side
    nb, nx, no = 64, 120, 9
    st = Step(ctx, nb, nx, no)
    hd = torch.empty(nb * nb, dtype=torch.float64).pin_memory(); hct = torch.empty(nb * no, dtype=torch.float64).pin_memory()
    hj = torch.empty(nb * nb, dtype=torch.float64).pin_memory(); hk = torch.empty(nb * nb, dtype=torch.float64).pin_memory()
    p = lambda t: C.c_void_p(t.data_ptr())

    def iteration():
        check(lib.rb_memcpy_h2d(ctx.h, p(st.dm), p(hd), hd.numel() * 8), "h2d")
        check(lib.rb_memcpy_h2d(ctx.h, p(st.ct), p(hct), hct.numel() * 8), "h2d")
        st.sh.dp(st.dm, out=st.d); st.sh.j(st.d, out=st.j); st.sh.k(st.ct, no, out=st.k)
        check(lib.rb_memcpy_d2h(ctx.h, p(hj), p(st.j), hj.numel() * 8), "d2h")
        check(lib.rb_memcpy_d2h(ctx.h, p(hk), p(st.k), hk.numel() * 8), "d2h")

    rng = np.random.default_rng(4)
    hd.copy_(torch.from_numpy(rng.standard_normal(nb * nb))); hct.copy_(torch.from_numpy(rng.standard_normal(nb * no)))
    iteration(); torch.cuda.current_stream().synchronize()
    with ctx.record() as rec:
        iteration()
    for it in range(3):
        hd.copy_(torch.from_numpy(rng.standard_normal(nb * nb))); hct.copy_(torch.from_numpy(rng.standard_normal(nb * no)))
        iteration(); torch.cuda.current_stream().synchronize()
        wj, wk = hj.clone(), hk.clone()
        hj.zero_(); hk.zero_()
        rec.graph.launch(); torch.cuda.current_stream().synchronize()
        assert torch.equal(hj, wj) and torch.equal(hk, wk), it
    rec.graph.close()


def test_single_pass_dp_j_can_be_recorded(side, monkeypatch):
    """rb_ri_dp_j's cooperative launch (and the sentinel fill of its exchange array) inside a recording: replay == direct call"""
    monkeypatch.setenv("REST_B200_DPJ_FUSED", "1")
    ctx = side
    st = Step(ctx, 128, 60, 8)
    d2, j2 = ctx.empty(60), ctx.empty(128 * 128)

    def seq():
        st.sh.dp_j(st.dm, out_d=d2, out_j=j2, reduce=False)
        st.sh.k(st.ct, st.no, out=st.k)
    seq()
    want = [d2.clone(), j2.clone(), st.k.clone()]
    with ctx.record() as rec:
        seq()
    for _ in range(3):
        d2.zero_(); j2.zero_(); st.k.zero_()
        rec.graph.launch()
        torch.cuda.current_stream().synchronize()
        assert torch.equal(d2, want[0]) and torch.equal(j2, want[1]) and torch.equal(st.k, want[2])
    rec.graph.close()
