"""GPU parity: the sm_100a path (through the C ABI) against the oracle on the same
fp32 state.  Integer outputs bit-exact; accelerations within 1e-5 relative
(BASELINE.json north_star); positions/velocities after one step accordingly.
"""
import os

import numpy as np
import pytest

from conftest import golden_files
from oracle import oracle as orc
from parity_util import acc_errors, f32_state, oracle_sim, random_state

pytestmark = pytest.mark.gpu

ACC_RTOL = 1e-5  # single-step accelerations (north_star)

FILES = golden_files()
IDS = [os.path.basename(f)[:-4] for f in FILES]


@pytest.fixture(scope="module")
def gof():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import game_of_flies_b200.integrator as integ
    import game_of_flies_b200.acceleration_updater as au
    import game_of_flies_b200.device as dev
    import game_of_flies_b200.engine as eng
    import game_of_flies_b200._lib as lib
    class NS: pass
    ns = NS()
    ns.torch, ns.integ, ns.au, ns.dev, ns.eng, ns.lib = torch, integ, au, dev, eng, lib
    return ns


def golden_state(path, step):
    g = np.load(path)
    st = dict(pos=f32_state(g[f"s{step}_in_pos"]), vel=f32_state(g[f"s{step}_in_vel"]),
              acc=f32_state(g[f"s{step}_in_acc"]), types=g["particle_type_index_array"].astype(np.int32),
              parameter_matrix=g["parameter_matrix"], limits=g["limits"], r_max=float(g["r_max"]),
              n=int(g["num_particles"]), T=int(g["num_types"]))
    ts = None if np.isnan(g["timestep"]) else float(g["timestep"])
    return st, ts, int(g["steps"])


def run_setup_bins(gof, pos, grid_dict):
    n = pos.shape[1]
    d = [gof.dev.to_device(pos[k]) for k in range(3)]
    i32 = gof.torch.int32
    pb, off, idx = (gof.dev.device_array(n, i32) for _ in range(3))
    cnt, st = (gof.dev.device_array(grid_dict["num_bins"], i32) for _ in range(2))
    gd = grid_dict
    gof.integ.setup_bins(d[0], d[1], d[2], gd["num_bin_x"], gd["num_bin_y"], gd["num_bin_z"],
                         gd["bin_size_x"], gd["bin_size_y"], gd["bin_size_z"], gd["num_bins"], n,
                         pb, cnt, off, st, idx)
    return d, pb, cnt, off, st, idx


def check_bins_against_oracle(gof, st):
    o = oracle_sim(st)
    o.setup_bins()
    d, pb, cnt, off, stt, idx = run_setup_bins(gof, st["pos"], o.grid)
    assert np.array_equal(pb.copy_to_host(), o.particle_bins)
    assert np.array_equal(cnt.copy_to_host(), o.particle_bin_counts)
    assert np.array_equal(stt.copy_to_host(), o.particle_bin_starts)
    assert np.array_equal(off.copy_to_host(), o.bin_offsets)
    assert np.array_equal(idx.copy_to_host(), o.particle_indices)
    return o, (d, pb, cnt, off, stt, idx)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_setup_bins_bit_exact_on_golden_states(gof, path):
    st, _, steps = golden_state(path, 0)
    check_bins_against_oracle(gof, st)
    st, _, _ = golden_state(path, steps - 1)
    check_bins_against_oracle(gof, st)


@pytest.mark.parametrize("limits,n", [((100, 100, 0), 50_000), ((120, 90, 150), 60_000), ((64, 64, 64), 1)])
def test_setup_bins_bit_exact_random_and_boundary_positions(gof, limits, n):
    st = random_state([n], limits, seed=7)
    grid = orc.grid_from_limits(st["limits"], st["r_max"])
    # adversarial: coordinates on and next to every bin edge, zero, and just inside the box
    pos = st["pos"]
    k = 0
    for axis, (nb, bs, L) in enumerate(((grid["num_bin_x"], grid["bin_size_x"], limits[0]),
                                        (grid["num_bin_y"], grid["bin_size_y"], limits[1]),
                                        (grid["num_bin_z"], grid["bin_size_z"], limits[2]))):
        if L == 0:
            continue
        edges = np.arange(nb) * bs
        cand = np.concatenate([edges, np.nextafter(edges.astype(np.float32), np.float32(np.inf)).astype(np.float64),
                               np.nextafter(edges[1:].astype(np.float32), np.float32(-np.inf)).astype(np.float64),
                               [np.nextafter(np.float32(L), np.float32(0))]])
        cand = cand.astype(np.float32).astype(np.float64)
        cand = cand[(cand >= 0) & (cand < L)]
        m = min(len(cand), pos.shape[1] - k)
        pos[axis, k:k + m] = cand[:m]
        k += m
    st["pos"] = f32_state(pos)
    check_bins_against_oracle(gof, st)


def run_accelerator(gof, st, bins, wrap_mode, with_boid_outputs=True):
    d, pb, cnt, off, stt, idx = bins
    n, T = st["n"], st["T"]
    v = [gof.dev.to_device(st["vel"][k]) for k in range(3)]
    types = gof.dev.to_device(st["types"])
    f32, i32 = gof.torch.float32, gof.torch.int32
    acc = [gof.dev.device_array(n, f32) for _ in range(3)]
    if with_boid_outputs:
        boid = [gof.dev.device_array(n * T, f32) for _ in range(6)]
        bc = gof.dev.device_array(n * T, i32)
    else:
        boid, bc = [None] * 6, None
    L = st["limits"]
    gof.au.accelerator(d[0], d[1], d[2], v[0], v[1], v[2], L[0], L[1], L[2], st["r_max"], n,
                       st["parameter_matrix"], types, acc[0], acc[1], acc[2], T, *boid, bc, None,
                       pb, idx, stt, cnt, wrap_mode=wrap_mode)
    got = np.stack([a.copy_to_host() for a in acc])
    return got, boid, bc


def check_accelerator(gof, st, wrap_mode):
    o, bins = check_bins_against_oracle(gof, st)
    o.wrap_mode = wrap_mode
    o.accelerate()
    got, boid, bc = run_accelerator(gof, st, bins, wrap_mode)
    want = np.stack([o.acc_x, o.acc_y, o.acc_z])
    assert np.array_equal(bc.copy_to_host(), o.boid_counts), "boid neighbour counts must be exact"
    mx, rms = acc_errors(got, want)
    assert mx <= ACC_RTOL, f"acceleration error {mx:.3e} (rms {rms:.3e})"
    if np.any(st["parameter_matrix"][-1] == 1):
        for k in range(3):
            np.testing.assert_allclose(boid[3 + k].copy_to_host(), o.boid[3 + k], rtol=1e-4, atol=1e-3)
            np.testing.assert_allclose(boid[k].copy_to_host(), o.boid[k], rtol=1e-4, atol=1e-2)
    return mx, rms


@pytest.mark.parametrize("wrap_mode", [orc.WRAP_REFERENCE, orc.WRAP_MINIMAGE], ids=["refwrap", "minimage"])
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_accelerator_on_golden_states(gof, path, wrap_mode):
    _, _, steps = golden_state(path, 0)
    for s in (0, steps - 1):  # the last recorded state has non-zero velocities (alignment, projection)
        st, _, _ = golden_state(path, s)
        check_accelerator(gof, st, wrap_mode)


@pytest.mark.parametrize("wrap_mode", [orc.WRAP_REFERENCE, orc.WRAP_MINIMAGE], ids=["refwrap", "minimage"])
@pytest.mark.parametrize("kinds,counts,limits,speed", [
    ("clusters", [3000] * 3, (100, 100, 0), 0.0),
    ("clusters", [4000] * 5, (120, 110, 130), 3.0),
    ("boids", [5000] * 3, (60, 70, 80), 9.0),
    ("boids", [3000] * 2, (90, 90, 0), 12.0),
    ("mixed", [2500] * 6, (110, 110, 110), 8.0),
    ("mixed", [1200] * 8, (100, 100, 100), 8.0),
    # Ly > Lz: the band y > Lz - r_max of the reference's z-wrap quirk covers half of the y range
    ("clusters", [4000] * 5, (100, 160, 90), 3.0),
    ("mixed", [2500] * 6, (90, 150, 100), 8.0),
    # Ly < Lz - r_max: no particle is in the band
    ("clusters", [4000] * 4, (120, 80, 130), 3.0),
])
def test_accelerator_random_states(gof, kinds, counts, limits, speed, wrap_mode):
    st = random_state(counts, limits, seed=11, kinds=kinds, speed=speed)
    check_accelerator(gof, st, wrap_mode)


def engine_from_state(gof, st, wrap_mode=orc.WRAP_REFERENCE):
    e = gof.eng.ParticleEngine(st["n"], st["parameter_matrix"], st["limits"], st["r_max"], wrap_mode=wrap_mode)
    e.upload([st["pos"][k].astype(np.float32) for k in range(3)], st["types"],
             vel=[st["vel"][k].astype(np.float32) for k in range(3)],
             acc=[st["acc"][k].astype(np.float32) for k in range(3)])
    return e


def check_engine_step(gof, st, ts, wrap_mode=orc.WRAP_REFERENCE):
    o = oracle_sim(st, timestep=ts, wrap_mode=wrap_mode)
    used = o.core_step()
    e = engine_from_state(gof, st, wrap_mode)
    e.step(1, ts)
    out = e.download(pos=True, vel=True, acc=True)
    pb, cnt, stt, idx = e.read_bins()
    cells = o.grid["num_cells"]
    assert np.array_equal(pb, o.particle_bins)
    assert np.array_equal(cnt, o.particle_bin_counts[:cells])
    assert np.array_equal(stt, o.particle_bin_starts[:cells])
    assert np.array_equal(idx, o.particle_indices)
    assert e.last_timestep() == pytest.approx(used, rel=1e-6)
    mx, rms = acc_errors(out["acc"], np.stack([o.acc_x, o.acc_y, o.acc_z]))
    assert mx <= ACC_RTOL, f"acceleration error {mx:.3e} (rms {rms:.3e})"
    # one step of dt <= 0.1 turns an acceleration error e into e*dt in v and e*dt^2/2 in x,
    # on top of fp32 rounding of the state itself (ulp of a coordinate ~ 1e-5 at 100)
    a_scale = np.abs(np.stack([o.acc_x, o.acc_y, o.acc_z])).max() + 1.0
    v_tol = 2e-5 * a_scale * used + 1e-5
    x_tol = 4 * np.spacing(np.float32(st["limits"].max())) + 1e-5 * a_scale * used * used
    want_v = np.stack([o.vel_x, o.vel_y, o.vel_z])
    want_x = np.stack([o.pos_x, o.pos_y, o.pos_z])
    assert np.abs(out["vel"] - want_v).max() <= v_tol
    dx = np.abs(out["pos"] - want_x)
    # a particle that lands within rounding of the wall may or may not have wrapped
    L = st["limits"][:, None]
    dx = np.where(L > 0, np.minimum(dx, np.abs(L - dx)), dx)
    assert dx.max() <= x_tol
    e.close()
    return mx


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_engine_step_on_golden_states(gof, path):
    _, ts, steps = golden_state(path, 0)
    for s in (0, steps - 1):
        st, ts, _ = golden_state(path, s)
        check_engine_step(gof, st, ts)


@pytest.mark.parametrize("kinds,counts,limits,speed,ts", [
    ("clusters", [20000] * 5, (190, 190, 190), 2.0, None),   # BASELINE density 0.014 / unit^3
    ("clusters", [5000] * 3, (270, 270, 0), 2.0, 0.02),
    ("boids", [20000] * 3, (160, 160, 160), 9.0, None),
    ("mixed", [8000] * 8, (165, 165, 165), 6.0, None),
    ("clusters", [12000] * 5, (150, 230, 140), 2.0, None),   # the quirk's y band spans many rows of cells
])
def test_engine_step_random_states(gof, kinds, counts, limits, speed, ts):
    st = random_state(counts, limits, seed=23, kinds=kinds, speed=speed)
    check_engine_step(gof, st, ts)
    check_engine_step(gof, st, ts, wrap_mode=orc.WRAP_MINIMAGE)


@pytest.mark.parametrize("wrap_mode", [orc.WRAP_REFERENCE, orc.WRAP_MINIMAGE], ids=["refwrap", "minimage"])
@pytest.mark.parametrize("counts,limits", [
    ([900] * 3, (100, 100, 0)),          # 2D: rows split over 3 warps
    ([1500] * 5, (60, 75, 45)),          # 3D, 3 z layers: most particles on the literal z-wrap path
    ([4000] * 4, (130, 120, 140)),       # 3D, interior and edge layers
])
def test_split_force_kernel_is_bit_identical_to_the_one_thread_kernel(gof, counts, limits, wrap_mode):
    """Small Clusters systems spread a home particle's stencil rows over 3 or 9 warps
    (gof_set_force_split); the partial sums are folded in the one-thread kernel's order."""
    st = random_state(counts, limits, seed=77, kinds="clusters", speed=2.0)
    lib = gof.lib.load()
    outs = {}
    try:
        for split in (1, 3, 9, -1):
            gof.lib.check(lib.gof_set_force_split(split))
            e = engine_from_state(gof, st, wrap_mode)
            e.step(3)
            outs[split] = e.download(pos=True, vel=True, acc=True)
            e.close()
            # the reference-layout entry point (accelerations scattered to particle order)
            bins = run_setup_bins(gof, st["pos"], oracle_sim(st).grid)
            outs[split]["scatter"], _, _ = run_accelerator(gof, st, bins, wrap_mode, with_boid_outputs=False)
    finally:
        gof.lib.check(lib.gof_set_force_split(-1))
    assert lib.gof_set_force_split(5) != 0  # rejected
    for split in (3, 9, -1):
        for k in ("pos", "vel", "acc", "scatter"):
            assert np.array_equal(outs[1][k], outs[split][k]), f"split {split}: {k} differs"


@pytest.mark.parametrize("kinds,counts,limits", [
    ("clusters", [700, 700, 600], (100, 100, 0)),
    ("mixed", [900] * 4, (70, 70, 70)),
])
def test_graph_launched_steps_equal_kernel_launched_steps(gof, kinds, counts, limits):
    """gof_engine_step replays one captured CUDA graph per step; switching timestep mode,
    parameters or wrap mode must re-capture, and nothing may differ from plain launches."""
    st = random_state(counts, limits, seed=31, kinds=kinds, speed=3.0)
    outs = []
    for graph in (True, False):
        e = engine_from_state(gof, st)
        e.set_graph(graph)
        e.step(5)              # adaptive timestep
        e.step(3, 0.02)        # fixed
        e.step(1, 0.01)        # another value: baked into the graph, must be re-captured
        e.step(4)
        pm = st["parameter_matrix"].copy()
        pm[3, 0, 1] *= 1.5
        e.set_parameters(pm)
        e.step(3)
        e.wrap_mode = orc.WRAP_MINIMAGE
        e.step(2)
        outs.append(e.download(pos=True, vel=True, acc=True))
        outs[-1]["dt"] = e.last_timestep()
        e.close()
    for k in ("pos", "vel", "acc"):
        assert np.array_equal(outs[0][k], outs[1][k]), f"{k}: graph and plain launches differ"
    assert outs[0]["dt"] == outs[1]["dt"]


@pytest.mark.parametrize("wrap_mode", [orc.WRAP_REFERENCE, orc.WRAP_MINIMAGE], ids=["refwrap", "minimage"])
def test_clustered_state_parity(gof, wrap_mode):
    """Parity on an EVOLVED Clusters state: after 150 steps the particles have clumped (cells with
    twice the mean population and more, long candidate runs, many chunks per run in the culled
    whole-warp walk); one more step from that state must agree with the oracle like any other."""
    st = random_state([5000] * 4, (105, 95, 110), seed=17, kinds="clusters", speed=1.0)
    e = engine_from_state(gof, st, wrap_mode)
    e.step(150)
    out = e.download(pos=True, vel=True, acc=True)
    _, cnt, _, _ = e.read_bins()
    e.close()
    assert cnt.max() > 2 * cnt.mean(), "the run should have clumped"
    evolved = dict(st)
    for k in ("pos", "vel", "acc"):
        evolved[k] = out[k].astype(np.float64)
    check_engine_step(gof, evolved, None, wrap_mode=wrap_mode)
    check_accelerator(gof, evolved, wrap_mode)


def test_engine_is_deterministic_and_keeps_every_particle(gof):
    st = random_state([30000] * 5, (220, 220, 220), seed=5, kinds="clusters", speed=1.0)
    outs = []
    for _ in range(2):
        e = engine_from_state(gof, st)
        e.step(5)
        outs.append(e.download(pos=True, vel=True, acc=True))
        pb, cnt, stt, idx = e.read_bins()
        assert np.array_equal(np.sort(idx), np.arange(st["n"])), "ids must stay a permutation"
        assert cnt.sum() == st["n"]
        assert np.all(np.diff(pb[idx]) >= 0), "particle_indices must be sorted by cell"
        e.close()
    for k in ("pos", "vel", "acc"):
        assert np.array_equal(outs[0][k], outs[1][k]), f"{k} differs between identical runs"
    L = st["limits"][:, None]
    assert np.all(outs[0]["pos"] >= 0) and np.all(outs[0]["pos"] <= L)


def test_engine_trajectory_statistics_against_oracle(gof):
    """Trajectories are chaotic: after many steps compare distributions, not particles."""
    st = random_state([4000] * 3, (150, 150, 0), seed=3, kinds="clusters")
    o = oracle_sim(st, timestep=None, n_threads=orc.max_threads())
    e = engine_from_state(gof, st)
    steps = 60
    for _ in range(steps):
        o.core_step()
    e.step(steps)
    out = e.download(pos=True, vel=True)
    spd_g = np.linalg.norm(out["vel"], axis=0)
    spd_o = np.linalg.norm(np.stack([o.vel_x, o.vel_y, o.vel_z]), axis=0)
    assert spd_g.mean() == pytest.approx(spd_o.mean(), rel=0.05)
    assert np.percentile(spd_g, 90) == pytest.approx(np.percentile(spd_o, 90), rel=0.1)
    # cell occupancy histogram (clustering) agrees
    pb, cnt, _, _ = e.read_bins()
    o.setup_bins()
    co = o.particle_bin_counts[:o.grid["num_cells"]]
    assert np.std(cnt) == pytest.approx(np.std(co), rel=0.15)
    e.close()


def test_long_run_stays_sane(gof):
    """400 steps of a mixed Clusters+Boids box (clusters form, cells become very uneven): every
    particle stays in the box, finite, ids remain a permutation, speeds respect the integrator's
    clamp-then-half-kick bound, and the binning of the last step is consistent."""
    st = random_state([2500] * 6, (100, 100, 100), seed=41, kinds="mixed", speed=3.0)
    e = engine_from_state(gof, st)
    e.step(400)
    out = e.download(pos=True, vel=True, acc=True)
    for k in ("pos", "vel", "acc"):
        assert np.all(np.isfinite(out[k])), k
    L = st["limits"][:, None]
    assert np.all(out["pos"] >= 0) and np.all(out["pos"] <= L)
    pb, cnt, stt, idx = e.read_bins()
    assert cnt.sum() == st["n"] and np.array_equal(np.sort(idx), np.arange(st["n"]))
    assert np.array_equal(stt, np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    assert cnt.max() > 3 * cnt.mean(), "the run should have clustered (uneven cells)"
    # |v| <= 20 after step1; step2 adds a * dt / 2 with dt <= 0.4 / max|v| of the step before
    assert np.linalg.norm(out["vel"], axis=0).max() < 60.0
    e.close()


def test_simulation_class_drop_in(gof, tmp_path):
    """The reference's Simulation surface: update(), output, update_parameter_matrix, blind_run + record."""
    from game_of_flies_b200.simulation import Simulation
    sim = Simulation(np.array([300, 300]), np.array([200]), seed=434, limits=(100, 100, 100))
    p0 = np.stack([sim.pos_x, sim.pos_y, sim.pos_z]).copy()
    sim.update()
    assert sim.output.shape == (800, 3) and sim.output.dtype == np.float64
    assert np.abs(sim.output.T - p0).max() < 1.0  # one small step from the initial positions
    assert not np.array_equal(sim.output.T, p0)
    params = [6.0, 11.0, -9.0, 3.0, 0]
    sim.update_parameter_matrix(0, 1, params)
    assert np.array_equal(sim.parameter_matrix[:, 0, 1], params)
    sim.update()
    rec = str(tmp_path / "run")
    sim.blind_run(3, rec)
    frames = sorted(os.listdir(os.path.join(rec, "frames")))
    # like the reference (simulation.py:328-345): the initial frame 00000 is overwritten by the
    # first step's frame because `frame` is only advanced inside the loop
    assert frames == ["00000.npy", "00001.npy", "00002.npy"]
    state = np.load(os.path.join(rec, "state.npz"))
    assert int(state["num_particles"]) == 800 and state["type_array"].shape == (800,)
    f = np.load(os.path.join(rec, "frames", "00002.npy"))
    assert f.shape == (800, 3)


def test_recorder_pipeline_writes_the_same_frames_as_stepwise_updates(gof, tmp_path):
    """blind_run(record=...) streams frame k to pinned memory (and to disk) while step k+1
    computes; the files must equal what update() + output gives step by step, and be readable
    the way the reference's npy2vid.py reads them (state.npz keys, frames/*.npy)."""
    from game_of_flies_b200.simulation import Simulation
    args = (np.array([500, 400, 300]), np.array([0]))
    kw = dict(seed=954, limits=(90, 90, 90))
    rec = str(tmp_path / "rec")
    a = Simulation(*args, **kw)
    a.update()
    a.blind_run(7, rec)
    b = Simulation(*args, **kw)
    b.update()
    frames = []
    for _ in range(7):
        b.update()
        frames.append(b.output.copy())
    names = sorted(os.listdir(os.path.join(rec, "frames")))
    assert names == [f"{k:05}.npy" for k in range(7)]
    for k, nm in enumerate(names):
        assert np.array_equal(np.load(os.path.join(rec, "frames", nm)), frames[k]), nm
    state = np.load(os.path.join(rec, "state.npz"))
    assert set(state.files) >= {"type_array", "limits", "parameter_matrix", "num_particles", "seed"}
    assert np.array_equal(state["type_array"][: int(state["num_particles"])], a.particle_type_index_array)


@pytest.mark.parametrize("kinds", ["clusters", "mixed"])
@pytest.mark.parametrize("two_buffer_sets", [True, False])
def test_host_driven_loop_with_overlapped_transfers_equals_stepwise(gof, kinds, two_buffer_sets):
    """gof_engine_step_host (uploads on the caller's stream, downloads on the engine's own, the next
    call's uploads ordered behind the copies into the buffers they read) must give, step for step,
    what upload + step + download with a synchronisation after every call gives -- whether the loop
    alternates two sets of host buffers or reads and writes one."""
    import torch
    st = random_state([1500, 1200, 900, 800], (70.0, 60.0, 80.0), 77, kinds=kinds, speed=2.0)
    n = st["n"]
    start = [st[g][k].astype(np.float32) for g in ("pos", "vel", "acc") for k in range(3)]
    types = torch.from_numpy(st["types"].astype(np.int32)).pin_memory()
    # reference: synchronous calls, state carried by the host
    a = gof.eng.ParticleEngine(n, st["parameter_matrix"], st["limits"], st["r_max"])
    cur = [x.copy() for x in start]
    want = []
    for _ in range(6):
        a.upload(cur[0:3], st["types"], vel=cur[3:6], acc=cur[6:9])
        a.step(1)
        d = a.download(pos=True, vel=True, acc=True)
        cur = [d[g][k].copy() for g in ("pos", "vel", "acc") for k in range(3)]
        want.append([x.copy() for x in cur])
    a.close()
    # the pipelined loop
    b = gof.eng.ParticleEngine(n, st["parameter_matrix"], st["limits"], st["r_max"])
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    sets = [[pin(x) for x in start], [pin(np.zeros(n, np.float32)) for _ in range(9)]]
    got = []
    for k in range(6):
        src, dst = (sets[k & 1], sets[(k & 1) ^ 1]) if two_buffer_sets else (sets[0], sets[0])
        b.step_host(src, types, dst)
        if k in (2, 5):  # look at the host buffers only where the loop waits
            b.host_wait()
            got.append((k, [t.numpy().copy() for t in dst]))
    b.close()
    for k, arrs in got:
        for q in range(9):
            assert np.array_equal(arrs[q], want[k][q]), (k, q)


def test_interleaved_host_driven_systems_equal_their_own_loops(gof):
    """bench.py's `e2e.interleaved_systems`: several independent systems, each in its own loop of
    gof_engine_step_host calls on its own stream (the copies of one run under the kernels of
    another), must each come out exactly as when it is stepped alone with synchronous calls."""
    import torch
    states = [random_state([900, 700, 800], (60.0, 70.0, 50.0), 31 + s, kinds="clusters" if s != 1 else "mixed",
                           speed=2.0) for s in range(3)]
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    want = []
    for st in states:
        a = gof.eng.ParticleEngine(st["n"], st["parameter_matrix"], st["limits"], st["r_max"])
        cur = [st[g][k].astype(np.float32) for g in ("pos", "vel", "acc") for k in range(3)]
        for _ in range(5):
            a.upload(cur[0:3], st["types"], vel=cur[3:6], acc=cur[6:9])
            a.step(1)
            d = a.download(pos=True, vel=True, acc=True)
            cur = [d[g][k].copy() for g in ("pos", "vel", "acc") for k in range(3)]
        want.append(cur)
        a.close()
    systems = []
    for st in states:
        e = gof.eng.ParticleEngine(st["n"], st["parameter_matrix"], st["limits"], st["r_max"])
        start = [st[g][k].astype(np.float32) for g in ("pos", "vel", "acc") for k in range(3)]
        sets = [[pin(x) for x in start], [pin(np.zeros(st["n"], np.float32)) for _ in range(9)]]
        systems.append([e, torch.cuda.Stream(), sets, pin(st["types"].astype(np.int32))])
    torch.cuda.synchronize()
    for _ in range(5):
        for sy in systems:
            e, stream, sets, types = sy
            with torch.cuda.stream(stream):
                e.step_host(sets[0], types, sets[1])
            sy[2] = [sets[1], sets[0]]
    for s, (e, stream, sets, _) in enumerate(systems):
        e.host_wait()
        stream.synchronize()
        for q in range(9):
            assert np.array_equal(sets[0][q].numpy(), want[s][q]), (s, q)
        e.close()


def check_full_field(gof, st, wrap_mode, label):
    """Whole-system parity at a BASELINE size: every particle's bin, the bin tables, the sorted
    order and the whole acceleration field of the state, then one full step, against the oracle
    (all host threads: its float64 atomics make the summation order vary at the 1e-16 level)."""
    from parity_util import acc_error_unfloored
    threads = orc.max_threads()
    o = oracle_sim(st, wrap_mode=wrap_mode, n_threads=threads)
    o.setup_bins()
    o.accelerate()
    want = np.stack([o.acc_x, o.acc_y, o.acc_z])
    e = engine_from_state(gof, st, wrap_mode)
    e.accelerate()
    got = e.download(pos=False, acc=True)["acc"]
    pb, cnt, stt, idx = e.read_bins()
    cells = o.grid["num_cells"]
    assert np.array_equal(pb, o.particle_bins)
    assert np.array_equal(cnt, o.particle_bin_counts[:cells])
    assert np.array_equal(stt, o.particle_bin_starts[:cells])
    assert np.array_equal(idx, o.particle_indices)
    mx, rms = acc_errors(got, want)
    raw = acc_error_unfloored(got, want)
    print(f"[{label}] N={st['n']} acceleration error: max {mx:.3e} (rms {rms:.3e}) relative to "
          f"max(|a_i|, median|a|); unfloored max |da_i|/|a_i| {raw:.3e}")
    assert mx <= ACC_RTOL, f"{label}: acceleration error {mx:.3e} (rms {rms:.3e})"
    e.close()
    # one full step (adaptive timestep) from the same state
    o2 = oracle_sim(st, timestep=None, wrap_mode=wrap_mode, n_threads=threads)
    used = o2.core_step()
    e = engine_from_state(gof, st, wrap_mode)
    e.step(1)
    out = e.download(pos=True, vel=True, acc=True)
    assert e.last_timestep() == pytest.approx(used, rel=1e-6)
    mx2, _ = acc_errors(out["acc"], np.stack([o2.acc_x, o2.acc_y, o2.acc_z]))
    assert mx2 <= ACC_RTOL
    a_scale = np.abs(np.stack([o2.acc_x, o2.acc_y, o2.acc_z])).max() + 1.0
    assert np.abs(out["vel"] - np.stack([o2.vel_x, o2.vel_y, o2.vel_z])).max() <= 2e-5 * a_scale * used + 1e-5
    dx = np.abs(out["pos"] - np.stack([o2.pos_x, o2.pos_y, o2.pos_z]))
    L = st["limits"][:, None]
    dx = np.where(L > 0, np.minimum(dx, np.abs(L - dx)), dx)
    assert dx.max() <= 4 * np.spacing(np.float32(st["limits"].max())) + 1e-5 * a_scale * used * used
    e.close()
    return mx, raw


@pytest.mark.parametrize("wrap_mode", [orc.WRAP_REFERENCE, orc.WRAP_MINIMAGE], ids=["refwrap", "minimage"])
def test_full_size_clusters3d_1m_whole_field(gof, wrap_mode):
    """BASELINE.json configs[1] (Clusters 3D, 5 species, N=1M, density of the reference's 3D run):
    the WHOLE field and the bins against the oracle, both wrap modes, no particle left out."""
    n_per = 200_000
    L = float((5 * n_per / 0.014) ** (1.0 / 3.0))
    st = random_state([n_per] * 5, (L, L, L), seed=1, kinds="clusters", speed=2.0)
    check_full_field(gof, st, wrap_mode, "clusters3d_1m")


def test_full_size_boids3d_4m_whole_field(gof):
    """BASELINE.json configs[2] (Boids 3D, N=4M): whole field, bins, exact neighbour structure
    (a wrong neighbour count shows as an O(1) error in cohesion / alignment), one full step."""
    n_per = 4_000_000 // 3
    st = random_state([n_per] * 3, (660.0, 660.0, 660.0), seed=2, kinds="boids", speed=8.0)
    check_full_field(gof, st, orc.WRAP_REFERENCE, "boids3d_4m")


def test_full_size_mixed3d_8sp_1m_whole_field(gof):
    """BASELINE.json configs[3]'s species mix (Clusters+Boids, 8 species) at N=1M on one GPU."""
    n_per = 125_000
    L = float((8 * n_per / 0.014) ** (1.0 / 3.0))
    st = random_state([n_per] * 8, (L, L, L), seed=3, kinds="mixed", speed=6.0)
    check_full_field(gof, st, orc.WRAP_REFERENCE, "mixed3d_8sp_1m")


def test_full_size_properties_clusters3d_1m(gof):
    """BASELINE.json configs[1] (Clusters 3D, 5 species, N=1M): size-independent properties."""
    n_per = 200_000
    L = float((5 * n_per / 0.014) ** (1.0 / 3.0))
    st = random_state([n_per] * 5, (L, L, L), seed=1, kinds="clusters")
    e = engine_from_state(gof, st)
    e.step(2)
    pb, cnt, stt, idx = e.read_bins()
    n = st["n"]
    assert cnt.sum() == n
    assert np.array_equal(stt, np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    assert np.array_equal(np.sort(idx), np.arange(n))
    assert np.all(np.diff(pb[idx]) >= 0)
    # stable inside every cell: ids ascending wherever the cell does not change
    same = np.diff(pb[idx]) == 0
    assert np.all(np.diff(idx)[same] > 0)
    out = e.download(pos=True, vel=True, acc=True)
    assert np.all(np.isfinite(out["acc"])) and np.all(np.isfinite(out["pos"]))
    # a random sample of particles against a brute-force float64 evaluation of the same state
    e2 = engine_from_state(gof, st)
    e2.accelerate()
    acc = e2.download(pos=False, acc=True)["acc"]
    rng = np.random.default_rng(0)
    sample = rng.choice(n, 64, replace=False)
    pos = st["pos"]
    pm = st["parameter_matrix"]
    Lv = st["limits"]
    for i in sample:
        d = pos[:, i:i + 1] - pos
        d -= Lv[:, None] * np.round(d / Lv[:, None])
        r2 = (d * d).sum(0)
        m = (r2 > 1e-10) & (r2 < st["r_max"] ** 2)
        r = np.sqrt(r2[m])
        tj = st["types"][m]
        a0, a1, a2, a3 = (pm[k, st["types"][i], tj] for k in range(4))
        f = np.where(r < a0, -a2 / a0 * r + a2,
                     np.where(r < (a0 + a1) / 2, 2 * a3 / (a1 - a0) * (r - a0), 2 * a3 / (a1 - a0) * (a1 - r)))
        want = -(f / r * d[:, m]).sum(1)
        z_edge = pos[2, i] < st["r_max"] or pos[2, i] > Lv[2] - st["r_max"] or pos[1, i] > Lv[2] - st["r_max"]
        if z_edge:
            continue  # the reference's z-wrap quirk applies there; covered by the oracle tests
        assert np.linalg.norm(acc[:, i] - want) <= 1e-5 * max(np.linalg.norm(want), 10.0)
    e.close(); e2.close()


def test_full_size_properties_boids3d_4m(gof):
    """BASELINE.json configs[2] (Boids 3D, N=4M): size-independent properties after two steps (the
    oracle comparison of the same system is test_full_size_boids3d_4m_whole_field)."""
    n_per = 4_000_000 // 3
    st = random_state([n_per] * 3, (660.0, 660.0, 660.0), seed=2, kinds="boids", speed=8.0)
    e = engine_from_state(gof, st)
    e.step(2)
    pb, cnt, stt, idx = e.read_bins()
    n = st["n"]
    assert cnt.sum() == n and np.array_equal(np.sort(idx), np.arange(n))
    assert np.all(np.diff(pb[idx]) >= 0)
    same = np.diff(pb[idx]) == 0
    assert np.all(np.diff(idx)[same] > 0)
    out = e.download(pos=True, vel=True, acc=True)
    for k in ("pos", "vel", "acc"):
        assert np.all(np.isfinite(out[k]))
    assert np.all(out["pos"] >= 0) and np.all(out["pos"] <= st["limits"][:, None])
    e.close()


@pytest.mark.parametrize("kinds,T", [("clusters", 16), ("mixed", 16), ("boids", 16), ("mixed", 1), ("clusters", 1)])
def test_species_count_limits(gof, kinds, T):
    """GOF_MAX_TYPES = 16 species (the largest shared-memory footprint of every kernel kind) and a
    single species: one step and the acceleration field against the oracle."""
    if kinds == "mixed" and T == 1:
        pytest.skip("a mixed system needs two species")
    st = random_state([max(6000 // T, 40)] * T, (90, 100, 110), seed=13, kinds=kinds, speed=4.0)
    check_engine_step(gof, st, None)
    check_accelerator(gof, st, orc.WRAP_REFERENCE)


@pytest.mark.parametrize("n", [1, 2, 31, 33, 129])
@pytest.mark.parametrize("kinds", ["clusters", "mixed", "boids"])
def test_tiny_and_ragged_particle_counts(gof, kinds, n):
    """Fewer particles than a warp / a block, and counts that leave ragged last warps."""
    T = 2
    counts = [n - n // 2, n // 2] if n > 1 else [1, 0]
    st = random_state(counts, (45, 50, 55), seed=n, kinds=kinds, speed=3.0)
    check_engine_step(gof, st, None)
    check_engine_step(gof, st, 0.02, wrap_mode=orc.WRAP_MINIMAGE)


@pytest.mark.parametrize("kinds", ["clusters", "mixed"])
def test_everything_in_a_few_cells(gof, kinds):
    """A clump: 6000 particles inside two neighbouring cells of a large, otherwise empty box (runs
    of thousands of candidates through the survivor lists, empty cells everywhere else)."""
    st = random_state([1500] * 4, (200, 200, 200), seed=21, kinds=kinds, speed=2.0)
    r = st["r_max"]
    rng = np.random.default_rng(5)
    pos = np.stack([rng.random(st["n"]) * 1.9 * r + 3 * r, rng.random(st["n"]) * 0.9 * r + 3 * r,
                    rng.random(st["n"]) * 0.9 * r + 3 * r])
    st["pos"] = f32_state(pos)
    check_engine_step(gof, st, None)
    check_accelerator(gof, st, orc.WRAP_MINIMAGE)
