import csv, collections, os, re, subprocess, sys
pat=sys.argv[1]; sasscsv=sys.argv[2]; cubin=sys.argv[3]
out=subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout
addr2line={};cur=None;fn=None
for line in out.splitlines():
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',line)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    m=re.match(r"\.text\.(\S+):",line)
    if m: fn=m.group(1); continue
    m=re.match(r"\s+/\*([0-9a-f]{4,6})\*/",line)
    if fn and pat in fn and m: addr2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(sasscsv)))
hdr=rows[1];col={h:i for i,h in enumerate(hdr)};data=rows[2:]
base=int(data[0][0],16)
ie=col['Instructions Executed'];ni=col['stall_no_inst'];ns=col['# Samples']
helpers=('warp_ctx.cuh','sm_30_intrinsics.hpp','sm_32_intrinsics.hpp','device_atomic_functions.hpp','device_functions.h','math_functions.hpp')
B=int(sys.argv[4]) if len(sys.argv)>4 else 0x800
buckets=collections.OrderedDict()
for r in data:
    a=int(r[0],16)-base
    b=a//B
    d=buckets.setdefault(b,[0,0,0,0,collections.Counter()])
    d[0]+=1; d[1]+=int(r[ie]); d[2]+=int(r[ns]); d[3]+=int(r[ni])
    l=addr2line.get(a)
    if l and l[0] not in helpers and not (l[0]=='boxqp_warp.cuh' and l[1]<60): d[4][(l[0][:14],l[1]//10*10)]+=1
tot=sum(d[2] for d in buckets.values())
for b,d in buckets.items():
    if d[1]==0: continue
    top=", ".join(f"{k[0]}:{k[1]}" for k,_ in d[4].most_common(3))
    print(f"{b*B:6x} exec/instr {d[1]/d[0]/1e6:7.2f}M samples {100*d[2]/tot:5.2f}% no_inst {100*d[3]/max(d[2],1):4.0f}%  {top}")
