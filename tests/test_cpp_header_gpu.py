"""The C++ host mirror (include/astroburst/core.hpp) EXECUTED on the GPU: a C++17 program stacks
deterministic frames through core::stacking::stack_images and the result is compared bit for bit
with the CPU oracle on the same frames (test_cpp_header_cpu.py only compiles and links it)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cstdint>
#include "astroburst/core.hpp"
using namespace astroburst;
int main(int argc, char **argv) {
    const size_t n = 40, rows = 37, cols = 101;   // 40 frames: the deep path; ragged sizes
    std::vector<Image> frames;
    uint64_t s = 12345;
    for (size_t f = 0; f < n; ++f) {
        frames.emplace_back(rows, cols);
        for (float &v : frames.back().data) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            v = 1000.0f + (float)((s >> 40) & 0xffff) * (1.0f / 256.0f);          // 1000 .. 1256
            if (((s >> 20) & 0x3ff) == 7) v += 30000.0f;                           // outliers
        }
    }
    try {
        Context ctx;
        core::stacking::StackConfig cfg;
        cfg.align = false;
        core::stacking::StackResult r = core::stacking::stack_images(ctx, frames, cfg);
        FILE *fp = fopen(argv[1], "wb");
        fwrite(r.image.data.data(), 4, r.image.data.size(), fp);
        fwrite(&r.rejected_pixels, 8, 1, fp);
        fclose(fp);
        try {  // error behaviour: bail!("No images to stack")
            core::stacking::stack_images(ctx, {}, cfg);
            return 3;
        } catch (const Error &) {}
    } catch (const Error &e) {
        fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    return 0;
}
'''


def test_core_hpp_runs_on_the_gpu_and_matches_the_oracle(orc):
    from astroburst_b200 import build

    so = build.build()
    n, rows, cols = 40, 37, 101
    s = np.uint64(12345)
    vals = np.empty(n * rows * cols, np.float32)
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    with np.errstate(over="ignore"):
        for i in range(vals.size):
            s = s * a + c
            v = np.float32(1000.0) + np.float32(int((s >> np.uint64(40)) & np.uint64(0xffff))) * np.float32(1.0 / 256.0)
            if int((s >> np.uint64(20)) & np.uint64(0x3ff)) == 7:
                v = np.float32(v + np.float32(30000.0))
            vals[i] = v
    frames = list(vals.reshape(n, rows, cols))
    want, want_rej = orc.stack_images(frames, 3.0, 3.0, 5)
    with tempfile.TemporaryDirectory() as td:
        cpp, exe, out = os.path.join(td, "t.cpp"), os.path.join(td, "t"), os.path.join(td, "out.bin")
        open(cpp, "w").write(SRC)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), cpp, "-o", exe,
                               so, f"-Wl,-rpath,{os.path.dirname(so)}"])
        r = subprocess.run([exe, out], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        raw = open(out, "rb").read()
    got = np.frombuffer(raw[: rows * cols * 4], np.float32).reshape(rows, cols)
    rej = int(np.frombuffer(raw[rows * cols * 4:], np.uint64)[0])
    assert rej == want_rej and want_rej > 0
    assert np.array_equal(got, want)
