"""Per-source-line instruction and stall-sample shares of one kernel: joins an `ncu --page source --csv` export (SASS rows, in
program order) with `nvdisasm -g -c` of the same cubin (line info).  usage: ncu_lines.py <src.csv> <all.sass> <mangled-name-substring> <source-file> [top]"""
import collections, csv, re, sys
src_csv, sass, key, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and key in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if (l.startswith('//--------------------- .text.') or l.startswith('.text.')) and seq:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), (m.group(3) or '').split('/')[-1], int(m.group(4) or 0))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst = []
for r in rows[2:]:
    try:
        inst.append((int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])))
    except Exception:
        pass
assert len(inst) == len(seq), (len(inst), len(seq))
byl, st = collections.Counter(), collections.Counter()
for (cur, txt), (cnt, sm) in zip(seq, inst):
    f, ln, f2, ln2 = cur
    k = (f, ln) if f == srcfile.split('/')[-1] or not f2 else (f + '<-' + f2, ln2 if f2 == srcfile.split('/')[-1] else ln)
    byl[k] += cnt
    st[k] += sm
tot, ts = sum(byl.values()), sum(st.values())
text = open(srcfile).read().split('\n')
print('warp instructions %d, stall samples %d' % (tot, ts))
for (f, ln), c in byl.most_common(top):
    t = text[ln - 1].strip()[:90] if srcfile.split('/')[-1] in f and 0 < ln <= len(text) else ''
    print('%5.1f%% instr %5.1f%% stalls  %s:%d  %s' % (100 * c / tot, 100 * st[(f, ln)] / ts, f, ln, t))
