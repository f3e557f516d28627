#!/bin/bash
# usage: tools/sass_lines.sh <mangled kernel name> : static SASS instruction count per source line
cd "$(dirname "$0")/.."; cuobjdump -xelf all mpp_cnn_rs_object_detection_b200/libmpp_b200.so >/dev/null 2>&1
nvdisasm -g mpp_b200.sm_100a.cubin 2>/dev/null | python3 -c "
import sys,re,collections
cur=None; cnt=collections.Counter(); fn=None
target=sys.argv[1]
for l in sys.stdin:
    m=re.match(r'\s*\.text\.(\S+):',l)
    if m: fn=m.group(1); continue
    if fn!=target: continue
    m=re.search(r'//## File \"([^\"]+)\", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]+\*/',l) and cur: cnt[cur]+=1
tot=sum(cnt.values()); print('total',tot)
byfile=collections.Counter()
for (f,ln),c in cnt.items(): byfile[f]+=c
print(byfile.most_common(8))
for (f,ln),c in cnt.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 40): print(c,f,ln)
" "$@"
rm -f *.cubin
