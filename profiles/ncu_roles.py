import csv,collections,sys
"""Stall samples, instruction counts and mbarrier-wait samples per warp role of the warp-specialised stage kernel (the
roles are delimited by their setmaxnreg instructions). usage: ncu -i report.ncu-rep --page source --csv > src.csv;
python profiles/ncu_roles.py src.csv [role whose top instructions to list]   (first kernel of the report only)"""
rows=list(csv.reader(open(sys.argv[1])))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
start=starts[0]
hdr=rows[start+1]; data=rows[start+2:(starts[1]-1 if len(starts)>1 else len(rows))]
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
idx={h:hdr.index(h) for h in stalls}
isrc=hdr.index('Source'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
def num(x):
    try: return int(x)
    except: return 0
marks=[i for i,r in enumerate(data) if len(r)>isrc and 'USETMAXREG' in r[isrc]]
regions={'pre':(0,marks[0]),'mma':(marks[0],marks[1]),'front':(marks[1],marks[2]),'back':(marks[2],len(data))}
for name,(a,b) in regions.items():
    c=collections.Counter(); n=0; ex=0
    for r in data[a:b]:
        if len(r)<=max(idx.values()): continue
        for h in stalls: c[h]+=num(r[idx[h]])
        n+=num(r[isamp]); ex+=num(r[iex])
    print(name,'samples',n,'instr',ex,[(h.replace('stall_',''),v) for h,v in c.most_common(7)])
    if len(sys.argv)>2 and name==sys.argv[2]:
        top=sorted([(num(r[isamp]),i,r[isrc][:80]) for i,r in enumerate(data[a:b]) if len(r)>isamp],reverse=True)[:14]
        for t in top: print('   ',t)
print('--- spin samples (SYNCS.PHASECHK and following BRA) per region')
for name,(a,b) in regions.items():
    tot=0; per=[]
    for i in range(a,b):
        r=data[i]
        if len(r)>isrc and 'SYNCS.PHASECHK' in r[isrc]:
            sm_=num(r[isamp])
            # following few instrs until BRA
            j=i+1; 
            while j<b and 'BRA' not in data[j][isrc]: sm_+=num(data[j][isamp]); j+=1
            sm_+=num(data[j][isamp]) if j<b else 0
            per.append((i,sm_)); tot+=sm_
    print(name,tot,per)
