import re,sys
def load(p):
    d={}
    for l in open(p):
        m=re.match(r'\| (\d+) \| ([^|]+) \| ([^|]+) \| ([\d.]+) \| ([\d.]+) \|',l)
        if m: d[int(m.group(1))]=(m.group(2).strip(),m.group(3).strip(),float(m.group(4)))
    return d
files=sys.argv[1:]
ds=[load(f) for f in files]
a=ds[0]
groups={}
for i in a:
    sh=a[i][1]
    key=sh.split('@')[0].strip()+' @'+sh.split('@')[1].split()[0] + (' +res' if '+res' in sh else '') + (' +up' if '+up' in sh else '')+(' nchw' if 'nchw' in sh else '')+(' partial' if 'partial' in sh else '')
    g=groups.setdefault(key,[0]+[0.0]*len(ds)); g[0]+=1
    for k,d in enumerate(ds): g[1+k]+=d[i][2]
print('%-48s  n  '%'family'+'  '.join('%9s'%f.split('/')[-1][:9] for f in files))
for k,g in sorted(groups.items(), key=lambda kv:-kv[1][1]):
    print('%-48s x%2d '%(k,g[0])+'  '.join('%9.1f'%v for v in g[1:]))
print('%-48s     '%'total'+'  '.join('%9.1f'%sum(g[1+k] for g in groups.values()) for k in range(len(ds))))
