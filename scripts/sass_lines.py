#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (needs -lineinfo).

  python scripts/sass_lines.py <lib.so> <cubin-name-part> <kernel-name-part> [top]

Extracts the cubins with cuobjdump, disassembles with nvdisasm -g and counts the instructions
that follow each '//## File ..., line N' marker inside the chosen .text section.  Used to see
what a source change does to the code without a GPU."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    lib, cub, kern = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
    f = [x for x in os.listdir(tmp) if x.startswith(cub + '.')][0]
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, inside, cnt = None, False, collections.Counter()
    ops = collections.Counter()
    for l in out.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', l)
        if m:
            inside = kern in m.group(1)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
        if m and cur:
            cnt[cur] += 1
            ops[m.group(2).split('.')[0]] += 1
    print('total static instructions', sum(cnt.values()))
    for k, c in cnt.most_common(top):
        print('%-12s %5d  %d' % (k[0], k[1], c))
    print(' '.join('%s:%d' % kv for kv in ops.most_common(25)))


if __name__ == '__main__':
    main()
