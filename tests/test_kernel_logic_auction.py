"""Kernel LOGIC of the KM replacement (gh-icp_b200/csrc/ghicp_auction.cu: CSC build, k_auc_init / k_auc_phase_start, the
persistent cooperative forward and reverse auction kernels, the free-object price check) run on the CPU through the host
emulation shim: km_auction itself drives an emulated cooperative launch in which every block of a small grid runs
concurrently as fibers and grid.sync() is a rendezvous.  Checked against scipy's exact assignment, the oracle's Kuhn-Munkres
(restated from src/km.cpp) and the reference's golden vectors, without a GPU; the -m gpu tests repeat this on the device."""
import ctypes as C
import json
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_longlong)


@pytest.fixture(scope="module")
def emu(emu_harness_path):
    L = C.CDLL(emu_harness_path)
    L.emu_km_auction.argtypes = [C.c_int, C.c_int, lp, ip, dp, C.c_double, C.c_double, ip, ip, dp, ip, ip]
    return L


def csr_from_gain(G):
    """Candidate edges = strictly positive gains (gain = penalty - CD; ghicp_capi.cu builds the same list)."""
    N, M = G.shape
    rowptr = np.zeros(N + 1, np.int64)
    col, gain = [], []
    for i in range(N):
        js = np.nonzero(G[i] > 0.0)[0]
        col.extend(js.tolist()); gain.extend(G[i, js].tolist())
        rowptr[i + 1] = len(col)
    return rowptr, np.array(col + [0], np.int32), np.array(gain + [0.0], np.float64)


def auction(emu, G, eps, monkeypatch=None):
    N, M = G.shape
    rowptr, col, gain = csr_from_gain(G)
    owner = np.full(M, -7, np.int32); assign = np.full(N, -7, np.int32); price = np.zeros(M)
    rounds, phases = C.c_int(0), C.c_int(0)
    mg = max(float(G.max()) if G.size else 0.0, eps)
    rc = emu.emu_km_auction(N, M, rowptr.ctypes.data_as(lp), col.ctypes.data_as(ip), gain.ctypes.data_as(dp), eps, mg,
                            owner.ctypes.data_as(ip), assign.ctypes.data_as(ip), price.ctypes.data_as(dp), C.byref(rounds),
                            C.byref(phases))
    assert rc == 0
    return owner, assign, price, rounds.value, phases.value


def check_matching(G, owner, assign):
    N, M = G.shape
    for j in range(M):
        if owner[j] >= 0:
            assert 0 <= owner[j] < N and assign[owner[j]] == j
            assert G[owner[j], j] > 0.0, "matched along a non-edge"
    for i in range(N):
        if assign[i] >= 0:
            assert owner[assign[i]] == i
    return float(sum(G[owner[j], j] for j in range(M) if owner[j] >= 0))


def optimum(G):
    """Best total gain when every row/column may also stay unmatched (gain 0): assignment on the clipped matrix."""
    Gp = np.maximum(G, 0.0)
    r, c = linear_sum_assignment(-Gp)
    return float(Gp[r, c].sum())


@pytest.mark.parametrize("N,M,density,seed,small", [(48, 48, 1.0, 1, None), (40, 56, 0.5, 2, "0"), (56, 40, 0.5, 3, "8"),
                                                   (64, 64, 0.15, 4, None), (64, 64, 0.15, 4, "0")])
def test_auction_within_n_eps_of_the_optimum(emu, monkeypatch, N, M, density, seed, small):
    if small is not None:
        monkeypatch.setenv("GHICP_AUCTION_SMALL", small)
    rng = np.random.default_rng(seed)
    pen = 1.0
    CD = rng.random((N, M)) * (pen / density)
    G = pen - CD                                      # <= 0 where CD >= penalty: not a candidate
    eps = 1e-3
    owner, assign, price, rounds, phases = auction(emu, G, eps)
    got = check_matching(G, owner, assign)
    assert phases > 1                                  # dense graph: epsilon scaling ran
    assert optimum(G) - got <= max(N, M) * eps + 1e-12
    assert rounds > 0


def test_auction_epsilon_complementary_slackness(emu):
    """What the optimality bound rests on: every assigned person is within eps of its best (gain - price, or staying out),
    every unassigned person finds nothing worth more than eps, and the prices of the objects left free sum to at most the
    budget the driver leaves for them (else the reverse auction must have cleared them)."""
    rng = np.random.default_rng(11)
    N, M, pen, eps = 50, 70, 1.0, 2e-3
    G = pen - rng.random((N, M)) * 1.6
    owner, assign, price, _, _ = auction(emu, G, eps)
    check_matching(G, owner, assign)
    Gp = np.where(G > 0, G, -np.inf)
    best = np.maximum((Gp - price[None, :]).max(axis=1), 0.0)
    for i in range(N):
        if assign[i] >= 0:
            assert G[i, assign[i]] - price[assign[i]] >= best[i] - eps - 1e-12
        else:
            assert best[i] <= eps + 1e-12
    free = owner < 0
    assert price[free].sum() <= 0.5 * eps * max(N, M) + 1e-12
    assert np.all(price >= 0.0)


def test_auction_sparse_graph_takes_the_single_phase_path(emu):
    """A settled loop: about one candidate per keypoint -> one forward phase, no scaling, no reverse auction."""
    rng = np.random.default_rng(5)
    n = 96
    G = np.full((n, n), -1.0)
    perm = rng.permutation(n)
    G[np.arange(n), perm] = 0.2 + 0.8 * rng.random(n)
    extra = rng.integers(0, n, size=(n // 3, 2))
    G[extra[:, 0], extra[:, 1]] = 0.1 + 0.5 * rng.random(len(extra))
    owner, assign, price, rounds, phases = auction(emu, G, 1e-3)
    got = check_matching(G, owner, assign)
    assert phases == 1
    assert optimum(G) - got <= n * 1e-3 + 1e-12
    assert np.all(price[owner < 0] == 0.0)             # nobody bid for them


@pytest.mark.parametrize("small", [None, "0", "4"])
def test_auction_contested_objects_need_the_reverse_phase(emu, capfd, monkeypatch, small):
    """Many persons compete for few valuable objects and then settle elsewhere only if the leftovers' prices are cleared:
    N << M with a price war, so free objects end a forward phase with positive prices.  The result must still be optimal
    to n*eps (that is what the reverse auction is for).  `small` = GHICP_AUCTION_SMALL: None -> at this size CTA 0 iterates
    alone (the tail path), "0" -> every round is a grid round (all CTAs, grid barriers), "4" -> the hand-over between the two."""
    rng = np.random.default_rng(8)
    N, M, eps = 24, 72, 1e-3
    G = 0.05 + 0.1 * rng.random((N, M))
    G[:, :6] += 0.8                                    # six objects everybody wants
    monkeypatch.setenv("GHICP_AUCTION_DEBUG", "1")
    if small is not None:
        monkeypatch.setenv("GHICP_AUCTION_SMALL", small)
    owner, assign, price, rounds, phases = auction(emu, G, eps)
    err = capfd.readouterr().err
    got = check_matching(G, owner, assign)
    assert optimum(G) - got <= max(N, M) * eps + 1e-12
    assert (assign >= 0).all()
    assert "-> reverse yes" in err
    grid_rounds = [int(l.split("grid rounds ")[1].split(")")[0]) for l in err.splitlines() if "grid rounds" in l][-1]
    assert (grid_rounds == 0) if small is None else (grid_rounds == rounds if small == "0" else 0 < grid_rounds < rounds)
    # the three schedules run the same rounds and end in the same matching: the result does not depend on who executes a round
    ref = _CONTESTED.setdefault("ref", (rounds, owner.copy(), assign.copy(), price.copy()))
    assert rounds == ref[0] and np.array_equal(owner, ref[1]) and np.array_equal(assign, ref[2]) and np.array_equal(price, ref[3])


_CONTESTED = {}


def test_reverse_phase_stops_on_the_price_budget(emu, capfd, monkeypatch):
    """The reverse rounds end as soon as D = the prices of the objects still free fits (1 - f) * n * eps (the bound
    OPT - ours <= n * eps_last + D holds at every round boundary): fewer rounds than running the chains to their end
    (GHICP_AUCTION_NOCUT), and both results within n * eps of the optimum."""
    rng = np.random.default_rng(11)
    N, M, eps = 120, 150, 1e-2
    G = np.floor(40 * rng.random((N, M))) - 30.0 + 0.37      # integer-spaced gains, about a quarter of them candidates, masses of ties
    monkeypatch.setenv("GHICP_AUCTION_DEBUG", "1")
    monkeypatch.setenv("GHICP_AUCTION_SCALING", "1")
    owner, assign, price, rounds, _ = auction(emu, G, eps)
    err = capfd.readouterr().err
    got = check_matching(G, owner, assign)
    monkeypatch.setenv("GHICP_AUCTION_NOCUT", "1")
    owner2, assign2, price2, rounds2, _ = auction(emu, G, eps)
    got2 = check_matching(G, owner2, assign2)
    opt = optimum(G)
    assert opt - got <= max(N, M) * eps + 1e-9 and opt - got2 <= max(N, M) * eps + 1e-9
    assert rounds <= rounds2
    if "stopped on the D budget" in err:
        D_left = float(err.split("D left ")[-1].split(" ")[0])
        assert D_left <= 0.9 * eps * max(N, M) + 1e-9
        assert price[owner < 0].sum() <= 0.9 * eps * max(N, M) + 1e-6


def test_auction_empty_and_single_edge_graphs(emu):
    G = np.full((5, 7), -1.0)
    owner, assign, _, rounds, _ = auction(emu, G, 1e-2)
    assert (owner == -1).all() and (assign < 0).all()     # rows without candidates are marked -2, never bid
    G[3, 2] = 0.5
    owner, assign, _, _, _ = auction(emu, G, 1e-2)
    assert owner[2] == 3 and assign[3] == 2 and (owner >= 0).sum() == 1


def test_auction_equals_oracle_km_on_the_reference_goldens(emu, orc, scratch_cwd):
    """Goldens G1/G2 and the committed km_golden.json cases: same objective as the restated Kuhn-Munkres (within n*eps;
    the matchings themselves coincide whenever the optimum is unique by more than that)."""
    with open(os.path.join(GOLD, "km_golden.json")) as f:
        cases = json.load(f)["cases"]
    for case in cases[:6]:
        W = np.array(case["W"], dtype=np.float64)
        n = W.shape[0]
        pen = -float(W.min())
        G = W + pen                                     # gain = penalty + weight
        eps = float(case["eps"])
        m = orc.km_solve(W, eps, "port")               # match[y] = x
        km_gain = float(sum(G[m[y], y] for y in range(n)))
        owner, assign, _, _, _ = auction(emu, G, eps)
        got = check_matching(G, owner, assign)
        assert abs(km_gain - got) <= n * eps + 1e-9
        assert optimum(G) - got <= n * eps + 1e-9
