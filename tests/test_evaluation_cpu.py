"""Host logic of dronechase_b200/evaluation.py (apps/threatsense_runner/evaluation_2bt.py over a batch): episode rows in
(step, env) order, the first N kept, pandas-style mean / sample std, result files."""
import numpy as np

from dronechase_b200.evaluation import episodes_from_steps, summarise, write_results


def test_rows_order_cap_and_stats(tmp_path):
    rows = []
    info = np.zeros((5, 8), dtype=np.int32)
    info[:, 0] = [3, 1, 4, 1, 5]; info[:, 1] = [2, 7, 1, 8, 2]; info[:, 2] = [0, 1, 0, 2, 0]; info[:, 3] = 3; info[:, 7] = 90
    episodes_from_steps(np.array([0, 1, 0, 1, 0], bool), info, 10, rows, 4)
    episodes_from_steps(np.array([1, 0, 1, 0, 1], bool), info, 11, rows, 4)
    assert [(r["step"], r["env"]) for r in rows] == [(10, 1), (10, 3), (11, 0), (11, 2)]        # capped at 4
    assert [r["total_kills"] for r in rows] == [8, 9, 5, 5]
    raw, stats = summarise(rows)
    assert raw["loyalwingman_0"] == [1, 1, 3, 4] and raw["loyalwingman_1"] == [7, 8, 2, 1]
    assert abs(stats["mean"]["total_kills"] - 6.75) < 1e-12
    assert abs(stats["std"]["total_kills"] - np.std([8, 9, 5, 5], ddof=1)) < 1e-12
    files = write_results(str(tmp_path / "out" / "results_2bt.xlsx"), raw, stats)
    assert files and all((tmp_path / "out" / f.split("/")[-1]).exists() for f in files)
