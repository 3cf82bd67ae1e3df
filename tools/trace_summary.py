"""Summarises a chrome trace written by `bench.py --trace`: per-stream busy time, idle gaps on the main stream and
the kernels around the largest gaps.  Run here (no GPU needed) on the file brought back in gpurun_out/."""
import collections
import gzip
import json
import re
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        return json.load(f)["traceEvents"]


def main(path, top=25):
    ev = [e for e in load(path) if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    streams = collections.defaultdict(list)
    for e in ev:
        streams[e["args"].get("stream")].append(e)
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    print(f"{len(ev)} device events over {(t1 - t0) / 1e3:.2f} ms on {len(streams)} streams")
    main_s = max(streams, key=lambda s: sum(e["dur"] for e in streams[s]))
    for s, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
        busy = sum(e["dur"] for e in es)
        print(f"  stream {s}: {len(es)} events, busy {busy / 1e3:.2f} ms{'  <- main' if s == main_s else ''}")
    es = streams[main_s]
    gaps = []
    for a, b in zip(es, es[1:]):
        gaps.append((b["ts"] - (a["ts"] + a["dur"]), a, b))
    tot_gap = sum(max(g[0], 0) for g in gaps)
    print(f"main stream: span {(es[-1]['ts'] + es[-1]['dur'] - es[0]['ts']) / 1e3:.2f} ms, busy {sum(e['dur'] for e in es) / 1e3:.2f} ms, "
          f"idle {tot_gap / 1e3:.2f} ms in {len(gaps)} gaps (median {sorted(g[0] for g in gaps)[len(gaps) // 2]:.1f} us)")
    hist = collections.Counter()
    for g, _, _ in gaps:
        hist[min(int(max(g, 0)) // 2 * 2, 20)] += 1
    print("  gap histogram (us -> count):", dict(sorted(hist.items())))
    short = lambda n: re.sub(r"\(.*", "", n)[:70]
    print(f"largest {top} gaps:")
    for g, a, b in sorted(gaps, key=lambda x: -x[0])[:top]:
        print(f"  {g:8.1f} us  after {short(a['name'])} ({a['dur']:.0f} us)  before {short(b['name'])}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in es:
        k = short(e["name"])
        agg[k][0] += 1
        agg[k][1] += e["dur"]
    print("main-stream kernels by time:")
    for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {d / 1e3:8.3f} ms {n:5d}  {k}")
    # gap attributed to the kernel that FOLLOWS it (launch latency of that kernel)
    by_next = collections.defaultdict(lambda: [0, 0.0])
    for g, a, b in gaps:
        k = short(b["name"])
        by_next[k][0] += 1
        by_next[k][1] += max(g, 0)
    print("idle time before (by following kernel):")
    for k, (n, d) in sorted(by_next.items(), key=lambda kv: -kv[1][1])[:20]:
        print(f"  {d / 1e3:8.3f} ms {n:5d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
