import sys,csv
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; body=[r for r in rows[2:] if len(r)==len(rows[1]) and r[0].startswith("0x")]
ia=hdr.index('Address'); isrc=hdr.index('Source'); ie=hdr.index('Instructions Executed'); iss=hdr.index('# Samples')
tot=sum(int(r[ie]) for r in body); tots=sum(int(r[iss]) for r in body)
print('total inst',tot,'samples',tots)
# contiguous regions: print cumulative by chunk of 40 instrs with top opcode
import collections
op=collections.Counter(); ops=collections.Counter()
for r in body:
    o=r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
    op[o]+=int(r[ie]); ops[o]+=int(r[iss])
for o,c in op.most_common(28): print('%-14s inst %5.1f%%  samples %5.1f%%'%(o,100*c/tot,100*ops[o]/tots))
if len(sys.argv)>2:
    top=sorted(body,key=lambda r:-int(r[iss]))[:int(sys.argv[2])]
    for r in top: print(r[ia][-5:], r[isrc][:70], 'inst',r[ie],'samples',r[iss])
