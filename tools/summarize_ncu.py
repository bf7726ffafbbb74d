"""Summarise gpurun_out ncu artefacts into profiles/ (launch-list shares + key metrics per kernel)."""
import collections, csv, json, subprocess, sys
tag = sys.argv[1]
launches, rep = sys.argv[2], sys.argv[3]
out = [f"# {tag}: ncu summary\n"]
# launch list
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.OrderedDict(); tot = 0.0
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("vt::<unnamed>::", "")[:48]
    v = float(row["Metric Value"].replace(",", "")); tot += v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
out.append("## Launch list (one forward, `--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t/1e3:.1f} | {t/tot*100:.1f} % |")
out.append(f"| total | {sum(n for n,_ in agg.values())} | {tot/1e3:.1f} | |\n")
# full capture
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = [("gpu__time_duration.sum", "time"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
idx = [(hdr.index(k), lab) for k, lab in want if k in hdr]
ni = hdr.index("Kernel Name")
out.append("## Full capture (`--set full --clock-control none`), one row per captured launch\n")
out.append("| kernel | " + " | ".join(f"{lab} ({units[i]})" if units[i] else lab for i, lab in idx) + " |")
out.append("|---|" + "---|" * len(idx))
traffic = {}
prev = ""
for row in r[2:]:
    if len(row) <= ni: continue
    name = row[ni].split("(")[0].replace("void ", "").replace("unnamed>::", "")[:40]
    out.append(f"| `{name}` | " + " | ".join(row[i] for i, _ in idx) + " |")
    after_gather, prev = "patch_gather" in prev, name
    # the four GEMMs of ONE layer (QKV, out-proj, fc1, fc2): not the token-mode patch-embedding GEMM that follows
    # the gather, not the next layer's launches
    if "gemm2" in name and not after_gather and len(traffic.get("per_launch", [])) < 4:
        rd = float(row[hdr.index("dram__bytes_read.sum")]); wr = float(row[hdr.index("dram__bytes_write.sum")])
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
        traffic.setdefault("per_launch", []).append((rd + wr) * mult)
open(f"profiles/{tag}_summary.md", "w").write("\n".join(out) + "\n")
if traffic:
    t = traffic["per_launch"]
    json.dump({"source": f"profiles/{tag}_summary.md (ncu --set full, gemm2 launches of one layer)",
               "dram_bytes_per_launch": sum(t) / len(t), "per_launch": t}, open("profiles/gemm_dram_traffic.json", "w"), indent=1)
print("\n".join(out))
