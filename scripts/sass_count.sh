#!/usr/bin/env bash
# Per-kernel SASS instruction mix of an object file / .so:  scripts/sass_count.sh file [name-filter]
cuobjdump -sass "$1" | awk -v filt="${2:-.}" '
/Function :/ {name=$3}
/^ +\/\*[0-9a-f]+\*\/ / { cnt[name]++;
  if ($0 ~ /FHFMA/) fh[name]++; if ($0 ~ /FFMA|FMUL|FADD/) ff[name]++; if ($0 ~ /LDG/) ld[name]++;
  if ($0 ~ /RED|ATOM/) rd[name]++; if ($0 ~ /SHFL/) sh[name]++; if ($0 ~ /PRMT|SHF\.|LOP3/) bo[name]++;
  if ($0 ~ /LDS|STS/) sm[name]++; if ($0 ~ /BAR/) br[name]++ }
END { for (n in cnt) if (n ~ filt) printf "%6d total  FHFMA=%-4d FP32=%-4d LDG=%-3d RED/ATOM=%-3d SHFL=%-3d bitops=%-4d LDS/STS=%-3d BAR=%-2d %s\n", cnt[n], fh[n], ff[n], ld[n], rd[n], sh[n], bo[n], sm[n], br[n], n }' | sort -k9
