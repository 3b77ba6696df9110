"""Per-stage warp-stall statistics of advop_mma_kernel from an ncu report (--set full --import-source on): SASS samples
are attributed to the kernel-body stage of the last seen source line (nvdisasm -g line info of the built library).
usage: python tools/ncu_stage_stalls.py <report.ncu-rep>   (scratch files under /tmp)"""
import csv,re,sys,subprocess
rep=sys.argv[1]
subprocess.run(f"ncu -i {rep} --page source --csv 2>/dev/null > /tmp/_src.csv",shell=True)
subprocess.run("cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all /root/repo/neko-top_b200/libneko_top_b200.so >/dev/null 2>&1; nvdisasm -g -c /tmp/cub/advop.sm_100a.cubin > /tmp/cub/dis.txt 2>/dev/null",shell=True)
lines=open('/tmp/cub/dis.txt').read().split('\n')
start=[i for i,l in enumerate(lines) if '.text._ZN4b20016advop_mma_kernelENS_12AdvMmaParamsE' in l and l.startswith('//---')][0]
end=next(i for i in range(start+1,len(lines)) if lines[i].startswith('//---------------------'))
cur=None; inst=[]; stack=[]
for l in lines[start:end]:
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2)),m.group(3)); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: inst.append((int(m.group(1),16),cur,m.group(2)))
rows=list(csv.reader(open('/tmp/_src.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; data=rows[hi+1:]; ci={h:i for i,h in enumerate(hdr)}
assert len(data)==len(inst),(len(data),len(inst))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
# walk in order; segment by BAR / named markers: attribute inline-function instrs to the last seen kernel-body line
src=open('/root/repo/neko-top_b200/csrc/advop_mma_kernel.cuh').read().split('\n')
kstart=next(i+1 for i,l in enumerate(src) if 'advop_mma_kernel(const __grid_constant__' in l)
marks=[(i+1,l.strip()[:60]) for i,l in enumerate(src) if i+1>kstart and ('// ---- ' in l or 'point-wise stage on the tile' in l or 'adjoint velocity: T only' in l or 'epilogue inputs' in l or '---- T1 on the warp' in l or l.strip().startswith('for (int it = blockIdx.x'))]
marks=[(kstart,'prologue')]+marks+[(len(src)+1,'end')]
tot=0; bucket={}; last='prologue'
for r,(addr,loc,txt) in zip(data,inst):
    n=int(r[ci['# Samples']] or 0); tot+=n
    if loc and loc[0]=='advop_mma_kernel.cuh' and loc[1]>=kstart:
        for (a,nm),(b,_) in zip(marks,marks[1:]):
            if a<=loc[1]<b: last=nm; break
    d=bucket.setdefault(last,{'n':0,'dmma':0})
    d['n']+=n
    if 'DMMA' in txt: d['dmma']+=int(r[ci['Instructions Executed']] or 0)
    for h in stalls: d[h]=d.get(h,0)+int(r[ci[h]] or 0)
print('total',tot)
for k,d in bucket.items():
    top=sorted([(h[6:],v) for h,v in d.items() if h.startswith('stall_')],key=lambda x:-x[1])[:5]
    print('%5.1f%%'%(100*d['n']/tot),'dmma/el %5d'%(d['dmma']//4096),k,top)
