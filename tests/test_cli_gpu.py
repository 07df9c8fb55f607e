"""The multi_frame_sr-compatible command line (multi_frame_sr.cpp:122-209): file naming, printed lines, outputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cli_city_like_run(cuda_device, tmp_path, capsys):
    cv2 = pytest.importorskip("cv2")
    from multi_frame_super_resolution_b200 import cli
    rng = np.random.default_rng(0)
    base = cv2.GaussianBlur(rng.integers(0, 255, size=(300, 400, 3), dtype=np.uint8), (0, 0), 2.0)
    for i in range(5):                                  # the repo's 0-based numbering (img_000000..000004.png)
        cv2.imwrite(str(tmp_path / f"img_{i:06d}.png"), np.ascontiguousarray(base[10 + i:10 + i + 192, 20 + 2 * i:20 + 2 * i + 256]))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert cli.main(["tiles", "city", "2"]) == 0
        assert cli.main(["a", "b"]) == -1               # usage (multi_frame_sr.cpp:136-142)
        assert cli.main(["tiles", "nosuch", "1"]) == -1  # "wrong input" (:161)
    finally:
        os.chdir(cwd)
    out = capsys.readouterr().out
    assert " sec" in out and " FPS" in out and "output_megapixels_per_second" in out
    res = cv2.imread(str(tmp_path / "city_tiles_sr_result.png"))
    assert res is not None and res.shape == (384, 512, 3) and res.std() > 5
    assert (tmp_path / "city_tiles_sr2_result.png").exists()
