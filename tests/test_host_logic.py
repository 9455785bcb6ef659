"""CPU: host-side logic of the batch layer that needs no device — the Crank-Nicolson countdown schedule, the solver
configuration struct, workload generators."""
import ctypes

import numpy as np
import pytest

from fastamericanoptionpricing_b200 import _lib, batch, workloads as wl
from oracle import alo_oracle as o


def test_cn_schedule_vectorised_equals_scalar_countdown():
    """cn_dt_schedule_batch == the reference's float countdown (CrankNicolsonOptionSolver.py:36-45) option by option, including
    the ~1e-16-long last step that appears when T is not an exact multiple of max_dt in floating point."""
    g = np.random.default_rng(7)
    T = np.concatenate([g.uniform(0.01, 5.0, 200), [3.0, 1.0, 0.25, 0.0]])
    mdt = np.concatenate([T[:100] / g.integers(1, 400, 100), g.uniform(0.001, 1.0, 100), [3.0 / 300, 1.0 / 12, 1.0, 0.1]])
    dts, ns = batch.cn_dt_schedule_batch(T, mdt)
    extra = 0
    for i in range(len(T)):
        ref = batch.cn_dt_schedule(T[i], mdt[i])
        assert ns[i] == len(ref) and np.array_equal(dts[i, :ns[i]], np.array(ref)) and np.all(dts[i, ns[i]:] == 0)
        assert ref == o.cn_dt_schedule(T[i], mdt[i])
        if len(ref) and ref[-1] < 1e-9 * mdt[i]:
            extra += 1
    assert extra >= 1 and len(batch.cn_dt_schedule(3.0, 3.0 / 300)) == 301      # the survey's example of the extra step
    assert batch.cn_dt_schedule(0.0, 0.1) == []


def test_config_struct_round_trip():
    cfg = batch.AloConfig("B", 24, 64, 128, 1e-7, 77, use_derivative=True, eta=0.75, kernel=1)
    c = cfg.c(_lib.FLAG_WS_SCRATCH)
    assert (c.solver, c.use_derivative, c.n_colloc, c.n_quad, c.n_price_quad, c.max_iters, c.kernel, c.flags) == (1, 1, 24, 64, 128, 77, 1, 1)
    assert c.iter_tol == 1e-7 and c.eta == 0.75 and batch.AloConfig().c().flags == 0
    d = batch.AloConfig()          # the reference's constructor defaults (FastAmericanOptionSolverBase.py:19-23, 46-47)
    assert (d.solver, d.n, d.l, d.m, d.iter_tol, d.max_iters, d.use_derivative, d.eta) == ("A", 12, 24, 48, 1e-5, 200, False, 0.5)
    assert d.rule_l == d.rule_m == "gauss-legendre" and d.seed == "hybr"
    assert batch.AloConfig(seed="newton").c().flags == _lib.FLAG_SEED_NEWTON and c.tau_max is None
    with pytest.raises(ValueError):
        batch.AloConfig(seed="bisect").c()
    assert ctypes.sizeof(c) == 56


def test_stress_workload_blocks():
    b = wl.stress_batch(3000)
    assert np.all(b["S"][:500] < 0.31 * b["K"][:500]) and np.all(b["S"][500:1000] > 2.9 * b["K"][500:1000])
    assert np.all(b["T"][1000:1500] <= 1e-2) and np.array_equal(b["q"][1500:1800], b["r"][1500:1800])
    assert np.all(b["q"][1800:2000] == 0) and np.all(b["r"][2000:2100] == 1e-6)
    b2 = wl.stress_batch(3000)
    assert all(np.array_equal(b[k], b2[k]) for k in b)       # seeded


def test_qdplus_internal_helpers_match_reference_values():
    """QDplus.q_QD / q_QD_dot / c0 / c / b / dlogSdh / h / M / N (reference QDplusAmericanOptionSolver.py:127-218, same signatures).
    Expected values: QDplus(0.04, 0.02, 0.25, 100, type).compute_miscellaneous(0.75, 93.0) run on the unmodified reference
    (numpy 2.3.5 / scipy 1.18.1); the European legs it needs are fed in as the reference computed them (they run on the device)."""
    from fastamericanoptionpricing_b200 import QDplusAmericanOptionSolver as m
    exp = {m.OptionType.Put: dict(v_p=11.130889999414507, v_qQD=-6.403484382807304, v_dlogSdh=-98.30766711885849, v_c=1.8370779650429843,
                                  v_c0=0.19775560768558423, v_b=-5.249805960975274),
           m.OptionType.Call: dict(v_p=5.7017470276485085, v_qQD=6.763484382807303, v_dlogSdh=-208.32127378593694, v_c=-2.802821633276153,
                                   v_c0=-10.979745079482687, v_b=5.249805960975274)}
    for ot, e in exp.items():
        s = m.QDplus(0.04, 0.02, 0.25, 100.0, ot)
        tau, S = 0.75, 93.0
        s.v_d1, s.v_d2, s.v_p, s.v_theta = -0.15765446457194918, -0.3741608155180588, e["v_p"], -3.7341196656609483
        s.v_N, s.v_M, s.v_h = s.N(), s.M(), s.h(tau)
        assert (s.v_N, s.v_M) == (0.64, 1.28) and abs(s.v_h - 0.029554466451491845) < 1e-17
        s.v_qQD, s.v_qQDdot = s.q_QD(tau), s.q_QD_dot()
        assert abs(s.v_qQDdot - 111.29558010352328) < 1e-12
        s.v_dlogSdh = s.dlogSdh(tau, S)
        got = dict(v_qQD=s.v_qQD, v_dlogSdh=s.v_dlogSdh, v_c=s.c(tau, S), v_c0=s.c0(tau, S), v_b=s.b(tau, S))
        for k, v in got.items():
            assert abs(v - e[k]) <= 1e-12 * max(1.0, abs(e[k])), (ot, k, v, e[k])


def test_host_buffer_checks_without_gpu():
    """HostContext._host_in / _host_out (the guards in front of the host-buffer C call) are plain host logic."""
    hin, hout = batch.HostContext._host_in, batch.HostContext._host_out
    x = np.arange(6, dtype=np.float64)
    assert np.shares_memory(hin(x, 6, np.float64, "x"), x)                   # right layout: passed through, no copy
    y = hin(x[::2], 3, np.float64, "x")
    assert y.flags.c_contiguous and np.array_equal(y, [0, 2, 4])
    assert hin([1, 2, 3], 3, np.float64, "x").dtype == np.float64 and hin(np.float32(2.5), 4, np.float64, "x").tolist() == [2.5] * 4
    f = np.array([True, False, True])
    assert hin(f, 3, np.uint8, "is_put").tolist() == [1, 0, 1] and hin(f.astype(np.int64), 3, np.uint8, "is_put").dtype == np.uint8
    for bad in (lambda: hin(x, 5, np.float64, "x"), lambda: hout(np.empty(3, dtype=np.int64), (3,), np.int32, "iters"),
                lambda: hout(np.empty(6)[::2], (3,), np.float64, "price"), lambda: hout(np.empty(4), (3,), np.float64, "price"),
                lambda: hout([0.0] * 3, (3,), np.float64, "price")):
        with pytest.raises(ValueError):
            bad()
    assert hout(np.empty((3, 13)), (3, 13), np.float64, "B").shape == (3, 13)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, timed on the
    plain-C port; needs no GPU and no /root/reference (the Python cross-reference leg reports itself absent)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FAOP_REF_STEP_SECONDS="0.5", FAOP_REFERENCE="/nonexistent")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"], cwd=root, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "American options priced/sec (FP64)" and line["unit"] == "options/s"
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("C2")
