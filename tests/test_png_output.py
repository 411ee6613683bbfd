"""the output path. rt_write_png (the role of stbi_write_png, optimized.cu:862) encodes the scanlines in
parallel bands concatenated into one zlib stream; rt_png_writer_* encodes and writes frames in the background. Every file must
decode, with an independent decoder, to exactly the bytes handed in. Runs without a GPU."""
import os

import numpy as np
import pytest

import raytracinggpu_b200 as rt

PIL = pytest.importorskip("PIL.Image")


def images():
    rng = np.random.RandomState(5)
    yield "noise_ragged", rng.randint(0, 256, size=(131, 77, 3)).astype(np.uint8)       # incompressible, odd sizes
    g = np.zeros((600, 800, 3), np.uint8)
    g[..., 0] = np.arange(800)[None, :] % 256
    g[..., 1] = (np.arange(600)[:, None] * 3) % 256
    yield "gradient_800x600", g                                                          # BASELINE.json configs[0] size
    yield "one_row", rng.randint(0, 256, size=(1, 5, 3)).astype(np.uint8)
    yield "black_4k", np.zeros((2160, 3840, 3), np.uint8)                               # configs[2] size, many bands


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_write_png_round_trip(built, tmp_path, monkeypatch, threads):
    monkeypatch.setenv("RT_PNG_THREADS", threads)
    for name, img in images():
        p = str(tmp_path / (name + ".png"))
        rt.write_png(p, img)
        back = np.asarray(PIL.open(p).convert("RGB"))
        assert back.shape == img.shape and np.array_equal(back, img), (name, threads)


def test_async_writer_writes_every_frame(built, tmp_path):
    w = rt.PngWriter(threads=4, max_pending=2)
    rng = np.random.RandomState(9)
    frames = [rng.randint(0, 256, size=(90, 160, 3)).astype(np.uint8) for _ in range(7)]
    buf = np.empty_like(frames[0])
    for k, f in enumerate(frames):
        buf[...] = f  # the caller's buffer is reused right after submit() returns
        w.submit(str(tmp_path / ("f%03d.png" % k)), buf)
    w.wait()
    for k, f in enumerate(frames):
        assert np.array_equal(np.asarray(PIL.open(str(tmp_path / ("f%03d.png" % k))).convert("RGB")), f)
    w.submit(str(tmp_path / "no_such_dir" / "x.png"), frames[0])  # an error in the background surfaces at wait()
    with pytest.raises(rt.RtError) as e:
        w.wait()
    assert e.value.code == -4
    w.close()
    assert os.path.exists(str(tmp_path / "f006.png"))
