"""tools/ncu_groups.py REPORT.ncu-rep COUNTERS.json OUT.json [source-note]

Turns one `ncu --set full --clock-control none` capture of `python tools/bench_stages.py ...` into the per-stage-group
figures bench.py reads (profiles/r02_dram_traffic.json): measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum)
and executed warp instructions (smsp__inst_executed.sum) per UNIT of the group (records scanned by L2, hits, probes, bases,
fragments, mappings).  COUNTERS.json holds the integer counters of the captured run (the `counters {...}` line bench_stages
prints, plus "bases").  Run where ncu is installed."""
import csv, io, json, re, subprocess, sys

GROUP_OF = [(r"(?<![a-z_])sketch_kernel|zip_records|table_fill|links_kernel|head_flags|unique_scatter|block_link_max|dir_fill", "hp1_index_build", "bases"), (r"lookup_kernel", "hp2_lookup", "probes"),
            (r"frag_l1_kernel|frag_l1_warp_kernel|frag_classify|cand_compact", "hp2_hits_l1", "hits"),
            (r"l2_bounds_kernel|l2_events_kernel|l2_seq_kernel|l2_kernel", "hp2_l2", "records"),
            (r"sort_unique|frag_ref_|compact_sketch", "hp2_query_sketch", "fragments"), (r"rows_kernel|cgi_", "hp2_report_cgi", "mappings")]

rep, cj, outp = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else rep
ctr = json.load(open(cj))
units = {"records": ctr["n2"], "hits": ctr["hits"], "probes": ctr["sum_s"], "bases": ctr["bases"], "fragments": ctr["fragments"], "mappings": ctr["mappings"]}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, un = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "gpu__time_duration.sum")}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
groups = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    for pat, g, unit in GROUP_OF:
        if re.search(pat, name):
            e = groups.setdefault(g, {"unit": unit, "dram_bytes": 0.0, "warp_inst": 0.0, "ms": 0.0, "kernels": {}})
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[un[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[un[col["dram__bytes_write.sum"]]]
            wi = float(r[col["smsp__inst_executed.sum"]].replace(",", ""))
            ms = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[un[col["gpu__time_duration.sum"]]]
            e["dram_bytes"] += rd + wr; e["warp_inst"] += wi; e["ms"] += ms
            k = e["kernels"].setdefault(re.sub(r"\(.*", "", name)[:48], {"launches": 0, "dram_GB": 0.0, "warp_inst": 0.0, "ms": 0.0})
            k["launches"] += 1; k["dram_GB"] += (rd + wr) / 1e9; k["warp_inst"] += wi; k["ms"] += ms
            break
out = {"what": "measured DRAM bytes and executed warp instructions per unit of each stage group, from ONE ncu --set full --clock-control none capture; "
               "bench.py multiplies by the units of the timed run to fill roofline.traffic and roofline.int_pipe", "source": note, "counters": ctr, "groups": {}}
for g, e in groups.items():
    u = units[e["unit"]]
    out["groups"][g] = {"unit": e["unit"], "units_in_capture": u, "dram_bytes_per_unit": e["dram_bytes"] / u, "warp_inst_per_unit": e["warp_inst"] / u,
                        "capture_ms": e["ms"], "source": note, "kernels": e["kernels"]}
json.dump(out, open(outp, "w"), indent=1)
for g, e in out["groups"].items():
    print("%-18s %10.2f B/%s  %8.2f warp-inst/%s  %8.2f ms" % (g, e["dram_bytes_per_unit"], e["unit"][:-1], e["warp_inst_per_unit"], e["unit"][:-1], e["capture_ms"]))
