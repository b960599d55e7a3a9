#!/bin/bash
# Static SASS evidence of the tcgen05 / TMEM / TMA paths per kernel (no GPU needed): cuobjdump + a mnemonic count.
set -e
cd "$(dirname "$0")/.."
cuobjdump -sass unirec_b200/libunirec_b200.so > /tmp/unirec_sass.txt
python - <<'PY'
import collections, re, subprocess
pat = re.compile(r'\b(UTC[A-Z0-9]*MMA|UTCBAR|UTCATOMSWS|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|HMMA|UCGABAR_ARV|UCGABAR_WAIT|FFMA2|MUFU)\b')
cur, counts = None, collections.OrderedDict()
for line in open('/tmp/unirec_sass.txt'):
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
    elif cur:
        for t in pat.findall(line):
            counts[cur][t] += 1
for f, c in counts.items():
    if c:
        name = re.sub(r'\(.*', '', subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip())
        print(f"{name[:90]:92s} " + "  ".join(f"{k}={c[k]}" for k in sorted(c)))
PY
