"""Critical path of a captured training step: reads the Graphviz dump of the step graph (USF_GRAPH_DOT=... with
DataParallelTrainer) and a device-op timeline (TIMELINE_DUMP=... of scripts/train_timeline.py) for per-kernel average
durations; prints the longest dependency chain, grouped, and what the first node of every late chain waits for."""
import collections, re, subprocess, sys
dot, events = sys.argv[1], sys.argv[2]
def short(n):
    n = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("usf::", "").replace("at::native::", "")
    return n.split("(")[0][:70]
dur = collections.defaultdict(list)
for line in open(events):
    p = line.split(None, 3)
    dur[p[3].strip()[:40]].append(float(p[1]))
avg = {k: sum(v) / len(v) for k, v in dur.items()}
text = open(dot).read()
names, kinds = {}, {}
for m in re.finditer(r'"graph_1_node_(\d+)"\[[^\]]*?label="\{(\w+)\s*\n\| \{ID \| \d+ \(topoId: \d+\) \| ([^\n]*?)\\<\\<\\<', text):
    names[int(m.group(1))] = m.group(3)
    kinds[int(m.group(1))] = m.group(2)
for m in re.finditer(r'"graph_1_node_(\d+)"\[[^\]]*?label="\{(\w+)', text):
    kinds.setdefault(int(m.group(1)), m.group(2))
mangled = sorted(set(names.values()))
dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.split("\n")
dm = {a: short(b) for a, b in zip(mangled, dem)}
label = {i: dm[n] for i, n in names.items()}
for m in re.finditer(r'"graph_1_node_(\d+)"\[', text):
    kinds.setdefault(int(m.group(1)), "MEMCPY")
for i, k in kinds.items():
    label.setdefault(i, k)
pred = collections.defaultdict(list)
for m in re.finditer(r'"graph_1_node_(\d+)" -> "graph_1_node_(\d+)"', text):
    pred[int(m.group(2))].append(int(m.group(1)))
n = max(kinds) + 1
def d(i):
    return avg.get(label[i][:40], 2.0)
fin, via = {}, {}
for i in range(n):                      # node ids are in capture order: predecessors have smaller ids
    best, arg = 0.0, None
    for p in pred.get(i, []):
        if fin[p] > best:
            best, arg = fin[p], p
    fin[i], via[i] = best + d(i) + 1.2, arg       # +1.2 us per dependent launch inside a graph
end = max(fin, key=fin.get)
print(f"{n} nodes, {sum(len(v) for v in pred.values())} edges; critical path {fin[end]:.0f} us (sum of average kernel times + 1.2 us per edge)")
path = []
i = end
while i is not None:
    path.append(i); i = via[i]
path.reverse()
agg = collections.OrderedDict()
seg, last = [], None
for i in path:
    seg.append((i, label[i], fin[i]))
tot = collections.defaultdict(lambda: [0, 0.0])
for i in path:
    tot[label[i]][0] += 1; tot[label[i]][1] += d(i)
print("critical path by kernel:")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"   {k:70s} {c:4d} {t:8.1f} us")
if len(sys.argv) > 3:
    for i, l, f in seg:
        print(f"{i:5d} {f:8.1f} {l}")
if len(sys.argv) > 4:
    for q in sys.argv[4:]:
        q = int(q)
        print(q, label[q], "<-", [(p, label[p][:30]) for p in pred.get(q, [])])
