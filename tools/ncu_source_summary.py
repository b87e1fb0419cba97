import csv,sys
f=sys.argv[1]; which=int(sys.argv[2]) if len(sys.argv)>2 else 0
rows=list(csv.reader(open(f)))
hdr=None; sec=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={'name':r[1],'rows':[]}; sec.append(cur); continue
    if r and r[0]=="Address": hdr=r; continue
    if cur is not None and len(r)>5: cur['rows'].append(r)
print(len(sec),'sections')
s=sec[which]['rows']
ia=hdr.index('Warp Stall Sampling (All Samples)'); ie=hdr.index('Instructions Executed')
tot=sum(int(r[ia] or 0) for r in s); print('total samples',tot, 'insts', sum(int(r[ie] or 0) for r in s))
W=int(sys.argv[3]) if len(sys.argv)>3 else 40
for b in range(0,len(s),W):
    t=sum(int(r[ia] or 0) for r in s[b:b+W]); ex=sum(int(r[ie] or 0) for r in s[b:b+W])
    ops=set((r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]) for r in s[b:b+W] if r[1].strip())
    key=[o for o in ops if o.startswith(('UTC','LDTM','UBLKCP','SYNCS','BAR','ATOM','LDG','RED','UTMA','MEMBAR','NANOSLEEP','FMNMX','FFMA'))]
    print(b, t, f"{100*t/tot:.1f}%", ex, sorted(key))
top=sorted(range(len(s)), key=lambda i:-int(s[i][ia] or 0))[:25]
for i in sorted(top):
    r=s[i]
    stalls={hdr[j][6:]:r[j] for j in range(hdr.index('stall_barrier'),len(hdr)) if j<len(r) and r[j] not in ('0','') and 'Not Issued' not in hdr[j]}
    print(i, r[1][:60], r[ia], r[ie], stalls)
