import csv,sys
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr=rows[0]
ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[iid],{'k':r[ik][:26]})[r[im]]=r[iv]
tot={}
for i,v in d.items():
    print(v['k'], ' '.join("%s=%s"%(k.split('__')[-1][:18],x) for k,x in v.items() if k!='k'))
    tot[v['k']]=tot.get(v['k'],0)+float(v.get('gpu__time_duration.sum',0))
print({k:round(x/1e3,1) for k,x in tot.items()})
