"""Scene generators reproduce the reference's only scene (Application::prepareDamBreak, Application.cpp:162-192)."""
import numpy as np

from akuaengine_b200 import scenes

F = np.float32


def test_dam_break_readme_scene():
    p, bmin, bmax = scenes.dam_break(30)
    assert len(p) == 27000  # NUM_PARTICLES, Application.cpp:12
    assert np.array_equal(bmin, F([1.5, 0.0, 1.5])) and np.array_equal(bmax, F([4.5, 4.0, 4.5]))  # :14-15
    # x outermost, z innermost; position = minPos + i * spacing in float (:173-181)
    assert np.array_equal(p["position"][0], F([2.0, 1.0, 2.0]))
    assert np.array_equal(p["position"][1], F([2.0, 1.0, F(2.0) + F(1) * F(0.05)]))
    assert np.array_equal(p["position"][30], F([2.0, F(1.0) + F(0.05), 2.0]))
    assert np.array_equal(p["position"][-1], F([F(2.0) + F(29) * F(0.05), F(1.0) + F(29) * F(0.05), F(2.0) + F(29) * F(0.05)]))
    assert np.all(p["mass"] == 1.0) and np.all(p["velocity"] == 0.0) and np.all(p["size"] == 50.0)
    assert np.all(p["color"] == F([0, 0, 1, 1]))


def test_scaled_dam_break_and_tank():
    p, bmin, bmax = scenes.dam_break(100)
    assert len(p) == 1_000_000
    assert np.allclose(bmin, [5.0, 0.0, 5.0], atol=1e-5) and np.allclose(bmax, [15.0, 13.3333, 15.0], atol=1e-3)
    assert np.all(p["position"].min(0) > bmin) and np.all(p["position"].max(0) < bmax)
    q, bmin, bmax = scenes.tank(40, 20, 10)
    assert len(q) == 8000 and np.all(q["position"].max(0) < bmax)
    g = scenes.tank_gravity(15.0)
    assert abs(np.linalg.norm(g) - 9.8) < 1e-5


def test_clouds_are_deterministic_and_bounded():
    a, bmin, bmax = scenes.uniform_cloud(5000)
    b, _, _ = scenes.uniform_cloud(5000)
    assert np.array_equal(a["position"], b["position"])
    assert np.all(a["position"] >= bmin) and np.all(a["position"] <= bmax)
    c, bmin, bmax = scenes.clustered_cloud(5000)
    assert np.all(c["position"] >= bmin) and np.all(c["position"] <= bmax)
    # density: ~8 particles per h-cell on average
    side = bmax[0]
    assert abs(5000 / (side / 0.1) ** 3 - 8.0) < 0.01


def test_scene_descriptions_build_the_baseline_configs():
    """scenes_json/*.json name real scenes of the sizes BASELINE.json quotes; building one needs no GPU."""
    import json
    from pathlib import Path
    from akuaengine_b200.run import build_scene, gravity_at
    want = {"config1_dambreak_27k.json": 27_000, "config2_dambreak_1m.json": 1_000_000}
    root = Path(__file__).resolve().parents[1] / "scenes_json"
    names = sorted(p.name for p in root.glob("*.json"))
    assert set(want) <= set(names) and len(names) >= 4
    for name in names:
        desc = json.loads((root / name).read_text())
        assert desc["scene"] in ("dam_break", "tank", "uniform_cloud", "clustered_cloud")
        if name in want:
            particles, bmin, bmax = build_scene(desc)
            assert len(particles) == want[name]
            p = particles["position"]
            assert np.all(p >= bmin) and np.all(p <= bmax)
    # piecewise-constant gravity schedule: the last entry whose time has been reached wins
    sched = [[0.0, 0, -9.8, 0], [0.5, 2.5, -9.5, 0], [1.0, -2.5, -9.5, 0]]
    assert gravity_at(sched, 0.25, [0, -1, 0]) == [0, -9.8, 0]
    assert gravity_at(sched, 0.5, [0, -1, 0]) == [2.5, -9.5, 0]
    assert gravity_at(sched, 7.0, [0, -1, 0]) == [-2.5, -9.5, 0]
    assert gravity_at([], 1.0, [0, -1, 0]) == [0, -1, 0]
