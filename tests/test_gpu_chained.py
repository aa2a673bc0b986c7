"""Chained launches (B200): consecutive plan steps of a batch that fills the GPU are issued back to back, the next step's lattice
kernel starting -- as a programmatic dependent -- on the SMs the previous step's last work items leave idle.  The
guards that keep an early starter from writing what an earlier launch still writes or reads (sequence-number gate for the
materialised rows, shadow block + griddepcontrol.wait for the cost / flags volume, alternating work counters) only matter when
the INPUTS DIFFER from step to step -- the bench repeats one batch and would not see a mix-up -- so these tests alternate
between different ego batches, lattices' time_step_now and output buffers without a synchronisation in between, and compare
with one isolated, synchronised call per configuration."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

B = 384  # x 5 horizons = 1920 pairs >= 2 x 148 SMs: the launches are issued directly and chain


def _setup():
    from test_gpu_graph_stream import _scene
    return _scene(batch=2 * B)


def _bufs(grid, want_mat):
    from test_gpu_graph_stream import _dev_buffers
    return _dev_buffers(B, grid, want_mat)


def _snap(b):
    return {k: (None if v is None else v.clone()) for k, v in b.items()}


def _same(x, y, what):
    for k in x:
        if x[k] is None:
            continue
        a, c = x[k].cpu().numpy(), y[k].cpu().numpy()
        assert np.array_equal(a, c, equal_nan=True), "%s: %s differs" % (what, k)


def _step(eng, ego_t, grid, prm, b, s):
    eng.plan_grid_dev(ego_t, grid, prm, b["cost"], b["flags"], b["mat"], b["idx"], b["best"], b["meta"], b["rec"], grid.n_stride, stream=s)


@pytest.mark.parametrize("want_mat", [False, True])
def test_train_of_different_steps_into_one_buffer_set(want_mat):
    """K steps, inputs alternating, ONE output set, no synchronisation: what is left must be the last step's results."""
    import torch
    sc, eng, grid, mk = _setup()
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    egos = [torch.tensor(np.ascontiguousarray(sc.ego[i * B:(i + 1) * B]), dtype=torch.float64, device=dev) for i in range(2)]
    prms = [mk(0), mk(11)]
    # isolated references, one synchronised call each
    ref = []
    for i in range(2):
        r = _bufs(grid, want_mat)
        _step(eng, egos[i], grid, prms[i], r, s)
        torch.cuda.synchronize()
        ref.append(_snap(r))
    assert not np.array_equal(ref[0]["idx"].cpu().numpy(), ref[1]["idx"].cpu().numpy())  # the two steps do differ
    out = _bufs(grid, want_mat)
    for k_steps in (2, 5, 8):
        for k in range(k_steps):
            _step(eng, egos[k % 2], grid, prms[k % 2], out, s)
        torch.cuda.synchronize()
        _same(ref[(k_steps - 1) % 2], out, "train of %d steps" % k_steps)


def test_train_with_a_buffer_set_per_step():
    """K winner-only steps into K output sets, no synchronisation: every step's results must be its own."""
    import torch
    sc, eng, grid, mk = _setup()
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    egos = [torch.tensor(np.ascontiguousarray(sc.ego[i * B:(i + 1) * B]), dtype=torch.float64, device=dev) for i in range(2)]
    prms = [mk(0), mk(11)]
    ref = []
    for i in range(2):
        r = _bufs(grid, False)
        _step(eng, egos[i], grid, prms[i], r, s)
        torch.cuda.synchronize()
        ref.append(_snap(r))
    outs = [_bufs(grid, False) for _ in range(6)]
    for k, o in enumerate(outs):
        _step(eng, egos[k % 2], grid, prms[k % 2], o, s)
    torch.cuda.synchronize()
    for k, o in enumerate(outs):
        _same(ref[k % 2], o, "step %d" % k)


def test_lattice_launches_alone_chain_too():
    """fiss_eval_grid_dev back to back (no record kernel in between): materialised rows and the volume of the last launch."""
    import torch
    sc, eng, grid, mk = _setup()
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    egos = [torch.tensor(np.ascontiguousarray(sc.ego[i * B:(i + 1) * B]), dtype=torch.float64, device=dev) for i in range(2)]
    prm = mk(0)
    ref = []
    for i in range(2):
        r = _bufs(grid, True)
        eng.eval_grid_dev(egos[i], grid, prm, r["cost"], r["flags"], r["mat"], grid.n_stride, stream=s)
        torch.cuda.synchronize()
        ref.append(_snap({k: r[k] for k in ("cost", "flags", "mat")}))   # (the winners' buffers stay unwritten here)
    out = _bufs(grid, True)
    for k in range(7):
        eng.eval_grid_dev(egos[k % 2], grid, prm, out["cost"], out["flags"], out["mat"], grid.n_stride, stream=s)
    torch.cuda.synchronize()
    for key in ("cost", "flags", "mat"):
        assert np.array_equal(out[key].cpu().numpy(), ref[0][key].cpu().numpy(), equal_nan=True), key


@pytest.mark.parametrize("want_mat", [False, True])
def test_shared_volume_winners_per_step(want_mat):
    """The case the shadow block is for: the steps SHARE the cost / flags volume (and the materialised rows) but every step has
    its own winners / records -- step k's record kernel reads the volume while step k+1's early CTAs are already producing
    theirs.  Every step's winners and records must be its own.  (Built with -DFISS_TEST_NO_SHADOW this test fails.)"""
    import torch
    sc, eng, grid, mk = _setup()
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    egos = [torch.tensor(np.ascontiguousarray(sc.ego[i * B:(i + 1) * B]), dtype=torch.float64, device=dev) for i in range(2)]
    prms = [mk(0), mk(11)]
    ref = []
    for i in range(2):
        r = _bufs(grid, want_mat)
        _step(eng, egos[i], grid, prms[i], r, s)
        torch.cuda.synchronize()
        ref.append(_snap(r))
    shared = _bufs(grid, want_mat)
    outs = [_bufs(grid, False) for _ in range(12)]
    for rep in range(3):
        for k, o in enumerate(outs):
            eng.plan_grid_dev(egos[k % 2], grid, prms[k % 2], shared["cost"], shared["flags"], shared["mat"], o["idx"], o["best"],
                              o["meta"], o["rec"], grid.n_stride, stream=s)
        torch.cuda.synchronize()
        for k, o in enumerate(outs):
            for key in ("idx", "best", "meta", "rec"):
                assert np.array_equal(o[key].cpu().numpy(), ref[k % 2][key].cpu().numpy(), equal_nan=True), \
                    "round %d step %d: %s differs" % (rep, k, key)
