import re,csv,collections,sys,subprocess,os
rep=sys.argv[1]; kern=sys.argv[2]; nreads=float(sys.argv[3])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],stdout=subprocess.PIPE).stdout.decode()
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ci={h:i for i,h in enumerate(hdr)}
lines=open('/tmp/xelf/dis.txt').read().split('\n')
start=next(i for i,l in enumerate(lines) if l.startswith('.text.') and kern in l)
cur=None; a2l={}; stack=[]
for l in lines[start+1:]:
    if l.startswith('.text.'): break
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2)), m.group(3)); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',l)
    if m: a2l[int(m.group(1),16)]=cur
def toint(a): return int(a,16) if a.startswith('0x') else int(a)
base=toint(data[0][ci['Address']])
# function ranges in fq_trim.cuh etc from source
import bisect
def func_ranges(path):
    txt=open(path).read().split('\n'); out=[]
    for i,l in enumerate(txt):
        m=re.match(r'^(?:template.*\n)?(?:__device__|__global__|static|inline|__host__).*?\b(\w+)\s*\(',l)
        if m and not l.startswith(' '): out.append((i+1,m.group(1)))
    return out
fr={f:func_ranges('/root/repo/faqcs_b200/csrc/'+f) for f in ('fq_trim.cuh','fq_emit.cuh','fq_frame.cuh')}
def fn(ln):
    if not ln: return 'none'
    f,l=ln[0],ln[1]
    if f not in fr: return f
    r=fr[f]; k=bisect.bisect_right([x[0] for x in r], l)-1
    return f+':'+(r[k][1] if k>=0 else '?')
agg=collections.Counter(); tot=0; samp=collections.Counter()
for r in data:
    off=toint(r[ci['Address']])-base
    n=float(r[ci['Instructions Executed']] or 0); tot+=n
    k=fn(a2l.get(off)); agg[k]+=n; samp[k]+=float(r[ci['# Samples']] or 0)
print('total inst/read', tot/nreads)
ts=sum(samp.values())
for k,n in agg.most_common(30): print(f"{n/nreads:7.1f} {n/tot*100:5.1f}%  samples {samp[k]/ts*100:5.1f}%  {k}")
