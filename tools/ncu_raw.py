import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]
want=sys.argv[1:]
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:90])
    for i,h in enumerate(hdr):
        if any(w in h for w in want): print(f'{h:75s} {r[i]:>18s} {units[i]}')
