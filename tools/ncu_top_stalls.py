import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; idx={h:i for i,h in enumerate(hdr)}
def num(x):
    try: return int(x)
    except: return 0
data=[r for r in rows[hi+1:] if len(r)==len(hdr) and r[0]!='Address']
tot=sum(num(r[idx['# Samples']]) for r in data)
print('total samples',tot)
# annotate with index to give context
keys=['stall_barrier','stall_long_sb','stall_wait','stall_sleep','stall_short_sb','stall_lg','stall_mio','stall_membar','stall_math','stall_tex','stall_branch_resolving','stall_dispatch','stall_no_inst']
order=sorted(range(len(data)),key=lambda i:-num(data[i][idx['# Samples']]))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]
for i in order:
    r=data[i]; n=num(r[idx['# Samples']])
    st={k.replace('stall_',''):num(r[idx[k]]) for k in keys}
    st={k:v for k,v in st.items() if v>0.15*n}
    print(f"{100*n/tot:5.1f}%  #{i:5d} {r[idx['Source']][:80]:80s} {st}")
