#!/usr/bin/env python
"""Executed instructions and stall samples of the stack kernel aggregated by code region.
usage: tools/ncu_regions.py rep [npix]  (regions are found by function-name markers in the source)"""
import csv, re, subprocess, sys
rep = sys.argv[1]
npix = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 4096
src_path = "astroburst_b200/csrc/stack_kernels.cuh"
marks = []  # (line, name) from the current source
for i, line in enumerate(open(src_path), 1):
    m = re.match(r"^(?:static )?(?:__device__|__global__|template|struct|constexpr|stack_\w+\()", line)
    m2 = re.search(r"\b(ce|sort_network_impl|sort_network|div_small|next_pow2_above|sum_run_mem|sumsq_run_mem|"
                   r"clip_combine_sorted_mem|finish_dynamic|finish_static|phase_a_tail|phase_a_uniform|phase_a|"
                   r"drain_impl|drain_dynamic|drop_non_finite|process_group|stack_warp_kernel|stack_wtma_kernel|"
                   r"mbar_\w+|tma_load_\w+|smem_u32)\s*\(", line)
    if m2 and (line.startswith("__device__") or line.startswith("static __device__") or line.startswith("stack_")
               or line.startswith("__global__")):
        marks.append((i, m2.group(1)))
def region(fn, l):
    if fn != "stack_kernels.cuh":
        return "intrinsics"
    name = "header"
    for ln, nm in marks:
        if ln <= l + 1:
            name = nm
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, fn, reg = None, "", {}
for r in rows:
    if r and r[0] == "File Path":
        fn = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) > ie and r[0].isdigit() and r[ie].isdigit():
        if "(ce<ce_is_imad" in r[1]:
            continue  # inlining aggregate of the ce() lines
        a = reg.setdefault(region(fn, int(r[0])), [0.0, 0])
        a[0] += int(r[ie]) / (npix / 32)
        a[1] += int(r[ss]) if r[ss].isdigit() else 0
ti = sum(v[0] for v in reg.values())
ts = sum(v[1] for v in reg.values())
for k, v in sorted(reg.items(), key=lambda x: -x[1][1]):
    if v[0] > 0:
        print(f"{k:26s} inst/group {v[0]:7.1f} ({100*v[0]/ti:4.1f}%)  samples {v[1]:6d} ({100*v[1]/ts:4.1f}%)"
              f"  rel. cycles/inst {v[1]/ts*ti/v[0]:.2f}")
print(f"total inst/group {ti:.1f}, samples {ts}")
