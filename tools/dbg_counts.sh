for k in text records mixed runs; do echo "== $k"; ZLB_LIB_OVERRIDE=$PWD/variants/dbg.so timeout 120 python tools/kind_case.py $k 16 2>&1 | grep chunk | head -400 | python -c "
import sys,re
rows=[list(map(int,re.findall(r'\d+',l))) for l in sys.stdin]
import statistics as st
n=len(rows)
if n:
    s=[sum(r[i] for r in rows) for i in range(11)]
    print('chunks',n,'spec: srch %d steps %d coop %d tok %d | fix: srch %d steps %d coop %d tok %d | iters/warp avg %.0f max-avg %.0f'%(s[1]/n,s[2]/n,s[3]/n,s[4]/n,s[5]/n,s[6]/n,s[7]/n,s[8]/n,s[9]/n/32,s[10]/n))
"; done
