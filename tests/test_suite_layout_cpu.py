"""The GPU suite's run order (tests/conftest.py) name every file that holds GPU tests."""
import glob
import os
import re

import conftest


def test_every_gpu_test_file_has_a_place_in_the_run_order():
    here = os.path.dirname(os.path.abspath(__file__))
    with_gpu = sorted(os.path.basename(p) for p in glob.glob(os.path.join(here, "test_*.py"))
                      if re.search(r"pytest\.mark\.gpu|\"cuda\"", open(p).read()) and os.path.basename(p) != os.path.basename(__file__))
    missing = [f for f in with_gpu if f not in conftest._GPU_FILE_ORDER]
    assert not missing, "add to tests/conftest.py _GPU_FILE_ORDER: %s" % missing
    # no GPU test may hide behind an expected-failure or skip marker: the hardware suite is strict
    for f in conftest._GPU_FILE_ORDER:
        src = open(os.path.join(here, f)).read()
        assert "xfail" not in src, f
