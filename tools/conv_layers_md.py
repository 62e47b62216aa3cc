#!/usr/bin/env python
"""profiles/conv_layers.md from the per-layer JSONs of bench/conv_layers.py (library columns: the --sweep run that timed
cuDNN; our columns: the last --sweep --skip_lib run of the final kernels) and bench/stem_bench.py."""
import json
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"
lib = {r["layer"]: r for r in json.loads((P / "conv_layers_lib_r2.json").read_text())["rows"]}
ours = {r["layer"]: r for r in json.loads((P / "conv_layers_ours_r2.json").read_text())["rows"]}
stem = json.loads((P / "stem_bench_r2.json").read_text())


def f(v):
    return "-" if v is None or v != v else f"{v:.1f}"


def ratio(o, l):
    return "-" if o is None or l is None or o != o or l != l else f"{o / l:.2f}"


out = ["# ResNet-50 convolutions, layer by layer: cuDNN vs this framework's tcgen05 kernels", "",
       "Batch 32, bf16, channels_last, one B200.  Microseconds per launch, CUDA events around the replay of a CUDA graph of 40",
       "back-to-back launches rotating over >= 192 MB of operands (cold L2), best of 3 (`bench/conv_layers.py`; stem:",
       "`bench/stem_bench.py`).  `x` = how many times the shape occurs in the network.  Our columns are the best tiling of the final",
       "kernels (`--sweep`; JSON: `conv_layers_ours_r2.json`), the library columns cuDNN with `cudnn.benchmark` (`conv_layers_lib_r2.json`).",
       "Ratio < 1 = ours faster.", "",
       "| layer | x | fprop cuDNN | ours | ratio | dgrad cuDNN | ours | ratio | wgrad cuDNN | ours | ratio |", "|---|---|---|---|---|---|---|---|---|---|---|"]
tot = {"lf": 0.0, "of": 0.0, "ld": 0.0, "od": 0.0, "lw": 0.0, "ow": 0.0}
for name, l in lib.items():
    o = ours.get(name, {})
    lf, ld, lw = l["lib_fprop_us"], l["lib_dgrad_us"], l["lib_wgrad_us"]
    of, od, ow = o.get("ours_fprop_us"), o.get("ours_dgrad_us"), o.get("ours_wgrad_us")
    if name.startswith("stem"):
        lf, lw = stem["lib_fprop_us"], stem["lib_wgrad_us"]
        of = stem["ours_fprop_total_us (pack x + pack w + conv, resident filter)"]
        ow = stem["ours_wgrad_dedicated_total_us (GEMM + reduce + unpack)"]
        ld = od = None
    out.append(f"| {name} | {l['count']} | {f(lf)} | {f(of)} | {ratio(of, lf)} | {f(ld)} | {f(od)} | {ratio(od, ld)} | {f(lw)} | {f(ow)} | {ratio(ow, lw)} |")
    c = l["count"]
    for key, lv, ov in (("f", lf, of), ("d", ld, od), ("w", lw, ow)):
        if lv is not None and lv == lv:
            tot["l" + key] += lv * c
            tot["o" + key] += (ov if ov is not None and ov == ov else lv) * c
out += ["", f"Count-weighted sums over the network (library time where we have no kernel - the six stride-2 layers): fprop cuDNN {tot['lf']:.0f} us vs "
        f"{tot['of']:.0f} us with ours, dgrad {tot['ld']:.0f} vs {tot['od']:.0f}, wgrad {tot['lw']:.0f} vs {tot['ow']:.0f}.", "",
        "## What runs by default, and why", "",
        "* **Stem (7x7, stride 2, 3 -> 64): ours, forward and weight gradient** - 4.5x / 4.1x faster than the library's padded mma.sync",
        "  kernels; the training step goes from 4.97 ms to 4.72 ms (`step_stem_{lib,native}_r2.json`, same box, back to back).",
        "  Decomposition in `stem.md`.",
        "* **Stride-1 1x1 / 3x3 layers: library** (`B200DDP_CONV=auto`).  In isolation our kernels win two shapes (layer1 64 -> 64",
        "  forward, layer1 64 -> 256 forward) and are within 1.0-1.5x elsewhere for forward / data gradient, 1.4-3x behind on weight",
        "  gradient; inside the captured training step the all-native configuration measured 5.43-5.80 ms against 4.95-5.00 ms",
        "  (`round2_ablation` runs), so the default keeps cuDNN there and the kernels stay selectable (`B200DDP_CONV=native`,",
        "  `B200DDP_BLOCK_FUSE=1` for the fused bottleneck node with BatchNorm statistics / backward sums in the conv epilogues).",
        "* **Stride-2 3x3 / 1x1 (six layers): library** - no kernel of ours.", "",
        "Where the gap on the inner layers comes from (cycle counters of CTA 0, `bench/conv_phases.py`): the activation producer is bound",
        "by the TMA issue rate (one 128-row box per ~190 cycles per issuing thread), the 3x3 taps re-fetch their shifted patch",
        "(nine boxes per 64 channels), and cuDNN's 1x1 kernels are 2-CTA (`nvjet_*_2cta`) with 256-wide tiles.  Next steps: `cta_group::2`",
        "with multicast of the filter tile, split-K across a cluster for the 7x7 / 14x14 maps, and a wgrad path that skips the fp32",
        "partial round trip when one CTA owns the whole pixel range."]
(P / "conv_layers.md").write_text("\n".join(out) + "\n")
print(P / "conv_layers.md")
