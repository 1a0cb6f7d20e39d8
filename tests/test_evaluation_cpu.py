"""Host logic of dronechase_b200/evaluation.py (apps/threatsense_runner/evaluation_2bt.py over a batch): a fixed quota of
episodes per env (no length bias), rows in (episode, env) order, pandas-style mean / sample std, result files."""
import numpy as np

from dronechase_b200.evaluation import episodes_from_steps, select_rows, summarise, write_results


def test_rows_order_cap_and_stats(tmp_path):
    rows = []
    info = np.zeros((5, 8), dtype=np.int32)
    info[:, 0] = [3, 1, 4, 1, 5]; info[:, 1] = [2, 7, 1, 8, 2]; info[:, 2] = [0, 1, 0, 2, 0]; info[:, 3] = 3; info[:, 7] = 90
    counts = episodes_from_steps(np.array([0, 1, 0, 1, 0], bool), info, 10, rows, 1)
    counts = episodes_from_steps(np.array([1, 1, 1, 0, 1], bool), info, 11, rows, 1, counts=counts)   # env 1 again: over quota
    assert counts.tolist() == [1, 1, 1, 1, 1] and len(rows) == 5
    rows = select_rows(rows, 4)                                                                  # cut by env index
    assert [(r["step"], r["env"]) for r in rows] == [(11, 0), (10, 1), (11, 2), (10, 3)]
    rows = sorted(rows, key=lambda r: (r["step"], r["env"]))
    assert [r["total_kills"] for r in rows] == [8, 9, 5, 5]
    raw, stats = summarise(rows)
    assert raw["loyalwingman_0"] == [1, 1, 3, 4] and raw["loyalwingman_1"] == [7, 8, 2, 1]
    assert abs(stats["mean"]["total_kills"] - 6.75) < 1e-12
    assert abs(stats["std"]["total_kills"] - np.std([8, 9, 5, 5], ddof=1)) < 1e-12
    files = write_results(str(tmp_path / "out" / "results_2bt.xlsx"), raw, stats)
    assert files and all((tmp_path / "out" / f.split("/")[-1]).exists() for f in files)
