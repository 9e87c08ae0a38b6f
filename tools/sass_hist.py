"""SASS opcode histogram per kernel of libb200wave.so (evidence for profiles/: FFMA2, LDGSTS, UBLKCP, SYNCS ...).
    python tools/sass_hist.py [path/to/lib.so] > profiles/rNN_sass_hist.txt"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'pytorch_wavelets_b200', 'libb200wave.so')
out = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, text=True).stdout
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', kern).replace('b200w::fast::', '').replace('b200w::', '').replace('void ', '')
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        total[m.group(1)] += 1
keys = ['FFMA2', 'FFMA', 'LDGSTS', 'UBLKCP', 'SYNCS', 'UTMALDG', 'UTMASTG', 'LDS', 'STS', 'LDG', 'STG', 'SHFL', 'BAR', 'LDL', 'STL']
print('libb200wave.so: %d kernels; totals: %s' % (len(hist), ', '.join('%s %d' % (k, total[k]) for k in keys)))
print('%-64s %7s ' % ('kernel', 'instrs') + ' '.join('%7s' % k for k in keys))
for k, c in hist.items():
    print('%-64s %7d ' % (k[:64], sum(c.values())) + ' '.join('%7d' % c[x] for x in keys))
