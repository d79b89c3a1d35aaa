"""instruction count per source line range of one kernel (needs -lineinfo):
   cuobjdump -xelf all lib.so; nvdisasm -c -g X.cubin > nvd.txt; python tools/sass_hist.py nvd.txt k_mega4ILi2E"""
import re, sys, collections
path, fn = sys.argv[1], sys.argv[2]
line = None; hist = collections.Counter(); infn = False
for ln in open(path):
    if ln.startswith('//---------------------'):
        infn = fn in ln
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', ln): hist[line] += 1
tot = sum(hist.values()); print('total instr', tot, '=', tot * 16 // 1024, 'KB')
b = collections.Counter()
for (f, l), c in hist.items(): b[(f, l // 10 * 10)] += c
for k, c in sorted(b.items(), key=lambda kv: -kv[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]: print(k, c)
