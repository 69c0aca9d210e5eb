import csv, re, sys
sass_file, src_csv, kern = sys.argv[1], sys.argv[2], sys.argv[3]
# 1. parse nvdisasm: instruction index -> source line for the kernel section
lines=open(sass_file).read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.') and kern in l][0]
cur=None; insn_line=[]
for l in lines[start+1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if insn_line: break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/',l):
        insn_line.append(cur)
# 2. sass page rows
rows=list(csv.reader(open(src_csv)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
h=rows[hi]; si=h.index('# Samples'); ie=h.index('Instructions Executed'); wf=h.index('L1 Wavefronts Shared')
body=rows[hi+1:]
print("sass rows",len(body),"nvdisasm insns",len(insn_line))
agg={}
for k,r in enumerate(body):
    key=insn_line[k] if k<len(insn_line) else None
    a=agg.setdefault(key,[0,0,0])
    a[0]+=int(r[si]); a[1]+=int(r[ie]); a[2]+=int(r[wf] or 0)
tot=sum(a[0] for a in agg.values()); tw=sum(a[2] for a in agg.values()); ti=sum(a[1] for a in agg.values())
src={}
def getsrc(f,n):
    import glob
    if f not in src:
        c=glob.glob('/root/repo/mpi_parallel_multiscale_diffusion_fem_b200/csrc/'+f)
        src[f]=open(c[0]).read().split('\n') if c else []
    return src[f][n-1].strip()[:95] if n-1<len(src[f]) else ''
print("total samples %d, wavefronts %d, warp-insts %d"%(tot,tw,ti))
for key,a in sorted(agg.items(),key=lambda kv:-kv[1][0])[:int(sys.argv[4]) if len(sys.argv)>4 else 40]:
    f,n=key if key else ('?',0)
    print("%5.1f%% smp %5.1f%% inst %5.1f%% wf  %s:%d  %s"%(100*a[0]/tot,100*a[1]/ti,100*a[2]/max(tw,1),f,n,getsrc(f,n) if key else ''))
