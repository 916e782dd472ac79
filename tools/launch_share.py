"""ncu launch list (gpu__time_duration.sum, CSV) of `bench.py --steps 1 --warmup 3 --no-cpu-baseline` -> per-kernel share table of
the TIMED step (markdown on stdout).  usage: python tools/launch_share.py gpurun_out/launches_final.csv [plain_bench_line.json]"""
import collections, csv, json, re, sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, idc = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
recs = [(int(r[idc]), r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) == len(h)]
k1 = [i for i, r in enumerate(recs) if "feature_fuse_staged" in r[1]]
# K1 launches: 1 (centroid build) + 2 per step x (3 warm-ups + 1 timed) + the end-to-end leg; the timed step starts at K1 launch #7
a, b = k1[7], k1[9]
step = recs[a:b]
last = max(i for i, r in enumerate(step) if "k_score" in r[1])
step = step[:last + 1]


def short(n):
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")


agg = collections.OrderedDict()
for _, n, t in step:
    k = short(n)
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
conv1 = [step[i][2] for i in range(len(step) - 1) if short(step[i][1]).startswith("k_gemm_tc<0") and short(step[i + 1][1]).startswith("k_gemm_tc<1")]
gn = sum(v[1] for k, v in agg.items() if k.startswith("k_gemm_tc<1"))      # MODE 1 = conv2 + residual + GELU + GroupNorm
print(f"{len(recs)} launches in the list; timed step: {len(step)} launches, {tot / 1e6:.2f} ms summed.\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / tot * 100:.1f} % | {v[1] / v[0] / 1e3:.1f} |")
blk = sum(v[1] for k, v in agg.items() if k.startswith("k_tcn_block"))      # fused TemporalConvBlock launches (conv1 + conv2 + GroupNorm)
print(f"\nfused TemporalConvBlock launches: {blk / 1e6:.2f} ms; conv1 launches of the two-kernel blocks (the `<0,1>` launch right before each `<1,1>`): "
      f"{len(conv1)}, {sum(conv1) / 1e6:.2f} ms; conv2 + GroupNorm launches {gn / 1e6:.2f} ms; "
      f"dilated-conv work together {(blk + sum(conv1) + gn) / 1e6:.2f} ms = **{(blk + sum(conv1) + gn) / tot * 100:.1f} % of the step**.")
gem = sum(v[1] for k, v in agg.items() if k.startswith("k_gemm_tc") or k.startswith("k_tlayer_tail")) - sum(conv1) - gn
k1t = sum(v[1] for k, v in agg.items() if "feature_fuse" in k)
oth = tot - gem - sum(conv1) - gn - blk - k1t
print(f"other GEMMs {gem / tot * 100:.1f} %, K1 {k1t / tot * 100:.1f} %, other kernels {oth / tot * 100:.1f} %.")
if len(sys.argv) > 2:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    sh = d["roofline"]["share_of_step"]
    t = sum(sh.values())
    print(f"bench.py CUDA-event classes of the plain run of the same command: conv {sh['conv_gemm_ms'] / t * 100:.1f} %, other GEMMs "
          f"{sh['other_gemm_ms'] / t * 100:.1f} %, K1 {sh['feature_fuse_ms'] / t * 100:.1f} %, other kernels {sh['other_kernels_ms'] / t * 100:.1f} %.")
