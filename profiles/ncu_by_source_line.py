import re,csv,collections,sys,subprocess,os
rep=sys.argv[1]; kern=sys.argv[2]; srcfile=sys.argv[3]; nreads=float(sys.argv[4])
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],stdout=subprocess.PIPE).stdout.decode()
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct','launch__occupancy_limit','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','sm__throughput.avg.pct','gpu__dram_throughput.avg.pct','launch__grid_size','launch__block_size','dram__throughput.avg.pct']
for h,u,v in zip(hdr,units,vals):
    if any(h==w or h.startswith(w) for w in want) and 'per_second' not in h and not h.endswith('.max') and '.min' not in h:
        print(f"{h:80s} {v:>16s} {u}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],stdout=subprocess.PIPE).stdout.decode()
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ci={h:i for i,h in enumerate(hdr)}
# disasm
os.system('cd /tmp && rm -rf xelf && mkdir xelf && cd xelf && cuobjdump -xelf all /root/repo/faqcs_b200/libfaqcs_b200.so >/dev/null && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null')
lines=open('/tmp/xelf/dis.txt').read().split('\n')
start=next(i for i,l in enumerate(lines) if l.startswith('.text.') and kern in l)
cur=None; a2l={}
for l in lines[start+1:]:
    if l.startswith('.text.'): break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',l)
    if m: a2l[int(m.group(1),16)]=cur
def toint(a): return int(a,16) if a.startswith('0x') else int(a)
base=toint(data[0][ci['Address']])
agg=collections.Counter(); samp=collections.Counter(); tot=0
for r in data:
    off=toint(r[ci['Address']])-base
    n=float(r[ci['Instructions Executed']] or 0); tot+=n
    ln=a2l.get(off); agg[ln]+=n; samp[ln]+=float(r[ci['# Samples']] or 0)
srcs={}
print('total inst/unit', tot/nreads)
for ln,n in agg.most_common(int(sys.argv[5]) if len(sys.argv)>5 else 30):
    text=''
    if ln:
        f='/root/repo/faqcs_b200/csrc/'+ln[0]
        if os.path.exists(f):
            srcs.setdefault(f,open(f).read().split('\n')); text=srcs[f][ln[1]-1].strip()[:80]
    print(f"{n/nreads:7.1f} {n/tot*100:5.1f}% samp={samp[ln]:7.0f} {ln} {text}")
