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
        // the same frames as BITPIX 16 data units (BZERO 32768, big-endian), decoded on the device
        std::vector<std::vector<uint8_t>> raw(n);
        std::vector<core::stacking::FitsUnit> units(n);
        for (size_t f = 0; f < n; ++f) {
            raw[f].resize(rows * cols * 2);
            for (size_t i = 0; i < rows * cols; ++i) {
                const int q = (int)frames[f].data[i] - 32768;                      // truncating: any i16 will do
                const uint16_t u = (uint16_t)(int16_t)(q < -32768 ? -32768 : q > 32767 ? 32767 : q);
                raw[f][2 * i] = (uint8_t)(u >> 8);
                raw[f][2 * i + 1] = (uint8_t)(u & 0xff);
            }
            units[f].data = raw[f].data();
            units[f].bitpix = 16;
            units[f].bzero = 32768.0;
        }
        core::stacking::StackResult r16 = core::stacking::stack_fits_units(ctx, units, rows, cols, cfg);
        fwrite(r16.image.data.data(), 4, r16.image.data.size(), fp);
        fwrite(&r16.rejected_pixels, 8, 1, fp);
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
    nb = rows * cols * 4
    got = np.frombuffer(raw[:nb], np.float32).reshape(rows, cols)
    rej = int(np.frombuffer(raw[nb:nb + 8], np.uint64)[0])
    assert rej == want_rej and want_rej > 0
    assert np.array_equal(got, want)
    # stack_fits_units: truncated i16 samples + 32768, stacked
    q = [(np.clip(f.astype(np.int64) - 32768, -32768, 32767) + 32768).astype(np.float32) for f in frames]
    want16, want16_rej = orc.stack_images(q, 3.0, 3.0, 5)
    got16 = np.frombuffer(raw[nb + 8:2 * nb + 8], np.float32).reshape(rows, cols)
    rej16 = int(np.frombuffer(raw[2 * nb + 8:], np.uint64)[0])
    assert rej16 == want16_rej and np.array_equal(got16, want16)
