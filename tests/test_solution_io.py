"""solution.txt wire format (writeSolution.writeDisplacement:8-20) and the reference's shipped golden vector."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from mofem_b200 import solution_io


def _shipped():
    return dict(np.load(os.path.join(GOLDEN, "shipped_solution_AMOREBracket.npz")))


def test_round_trip_is_lossless_at_the_files_precision():
    g = _shipped()
    text = solution_io.format_solution(g["displacement"], g["energy"], str(g["name"]))
    d, e, name = solution_io.parse_solution(text)
    assert name == "AMOREBracket" and d.shape == (2222, 1) and e.shape == (1, 1)
    assert np.array_equal(d.reshape(-1), g["displacement"]) and e[0, 0] == g["energy"][0]
    assert text.endswith("The strain ENERGY is: 1.061718786130466e+00.")
    assert "  1110            2.015749752214377e-06           -2.540141808564546e-06\n" in text


@pytest.mark.skipif(not os.path.exists("/root/reference/solution.txt"), reason="reference checkout not present")
def test_writer_reproduces_the_shipped_file_byte_for_byte(tmp_path):
    g = _shipped()
    solution_io.write_solution(tmp_path / "solution.txt", g["displacement"].reshape(-1, 1), g["energy"].reshape(1, 1),
                               "AMOREBracket")
    assert (tmp_path / "solution.txt").read_text() == open("/root/reference/solution.txt").read()


def test_live_reference_run_agrees_with_the_shipped_golden_vector():
    # the fixture AMOREBracket.npz is the unmodified reference run in the build container
    _, d = load_golden("AMOREBracket")
    g = _shipped()
    du, de = solution_io.compare(d["displacement"], d["energy"], g["displacement"], g["energy"])
    assert du <= 1e-9 and de <= 1e-9


def test_parse_errors():
    with pytest.raises(ValueError):
        solution_io.parse_solution("nothing here")
    with pytest.raises(ValueError):
        solution_io.compare(np.zeros(4), None, np.ones(6))
