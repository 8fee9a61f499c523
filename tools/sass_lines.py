"""Static per-source-line SASS instruction counts of one kernel (nvdisasm -g -c output): an offline stand-in for
the per-line instruction profile -- which lines of the kernel loop the compiler turned into how many instructions."""
import re, sys, collections
path, func, src = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(path).read().split("\n")
text = open(src).read().split("\n")
infunc = False; cur = None; counts = collections.Counter(); order = []
for ln in lines:
    if ln.startswith(".text."):
        infunc = func in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", ln) and cur:
        counts[cur] += 1
tot = sum(counts.values())
print("total instructions", tot)
for (f, l), c in sorted(counts.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    t = text[l - 1].strip()[:110] if f == src.split("/")[-1] and l <= len(text) else ""
    print("%4d  %s:%d  %s" % (c, f, l, t))
