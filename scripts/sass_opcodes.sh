#!/bin/bash
# Scripted SASS opcode count of the shipped library: the mnemonics that prove which hardware paths the kernels use
# (fp64 tensor path DMMA, cp.async LDGSTS, tcgen05 UTCIMMA / TMEM LDTM / TMA UTMALDG / UTCBAR / SYNCS, dp4a IDP).
#   bash scripts/sass_opcodes.sh > profiles/r2_sass_opcodes.txt
cd "$(dirname "$0")/.."
SO=gaussian_process_optimization_b200/libgpb200.so
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | grep release | sed 's/.*release //'))"
cuobjdump -sass "$SO" > /tmp/gpb_sass.txt
echo "# functions: $(grep -c 'Function :' /tmp/gpb_sass.txt)"
for op in DMMA DFMA DADD DMUL MUFU LDGSTS UTCIMMA UTCHMMA UTCQMMA LDTM STTM UTMALDG UTMASTG UTCBAR SYNCS IDP IMAD PRMT REDUX ATOM RED; do
  printf "%-8s %s\n" "$op" "$(grep -cE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?$op[. ]" /tmp/gpb_sass.txt)"
done
echo "# per kernel (DMMA / UTCIMMA / LDTM / UTMALDG / LDGSTS):"
awk '/Function :/ {name=$3} /DMMA/ {d[name]++} /UTCIMMA/ {u[name]++} /LDTM/ {l[name]++} /UTMALDG/ {t[name]++} /LDGSTS/ {g[name]++}
     END {for (n in d) printf "%s DMMA=%d\n", n, d[n]; for (n in u) printf "%s UTCIMMA=%d LDTM=%d UTMALDG=%d\n", n, u[n], l[n], t[n]; for (n in g) printf "%s LDGSTS=%d\n", n, g[n]}' /tmp/gpb_sass.txt | c++filt | cut -c1-160 | sort
