"""The C++ host mirror (include/astroburst/core.hpp) and the C header compile as C++17 / C11 and
link against the built library (no GPU needed: nothing is executed)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_headers_compile_and_link():
    from astroburst_b200 import build

    so = build.build()
    with tempfile.TemporaryDirectory() as td:
        c_src = os.path.join(td, "t.c")
        open(c_src, "w").write('#include "astroburst_cuda.h"\nint main(void){return ab_abi_version()==AB_ABI_VERSION?0:1;}\n')
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               c_src, "-o", os.path.join(td, "tc"), so, f"-Wl,-rpath,{os.path.dirname(so)}"])
        assert subprocess.call([os.path.join(td, "tc")]) == 0
        cpp = os.path.join(td, "t.cpp")
        open(cpp, "w").write(
            '#include "astroburst/core.hpp"\n'
            "int main(){ using namespace astroburst; try { Context c; Image r(2, 2), g(2, 2), b(2, 2);\n"
            " core::imaging::ToneChain t; core::imaging::apply_tone_composite(c, r, g, b, t);\n"
            " core::preview::masked_stretch(c, r); } catch (const Error&) { return 0; }\n"
            " return 0; }\n")
        subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), cpp, "-o",
                               os.path.join(td, "tcpp"), so, f"-Wl,-rpath,{os.path.dirname(so)}"])
        assert subprocess.call([os.path.join(td, "tcpp")]) == 0
