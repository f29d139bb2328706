#!/bin/bash
# SASS listings of the two dominant kernels (instruction text only) -> profiles/.  Run after `make -C <pkg>/csrc`.
# usage: bash tools/dump_sass.sh [tag]
TAG=${1:-r2}
cd "$(dirname "$0")/.."
B=$(ls -d dh-aug*/build)
dump() {   # object file, mangled-name fragment, output name
  cuobjdump -sass "$B/$1" | awk -v pat="$2" '/Function : /{p=index($0,pat)>0} p' \
    | grep -vE "^\s+/\* 0x[0-9a-f]+ \*/\s*$" | sed -E 's#\s+/\* 0x[0-9a-f]+ \*/\s*$##' > "profiles/$3"
  echo "$3: $(grep -cE '^\s+/\*[0-9a-f]{4,5}\*/' profiles/$3) instructions;" \
       "$(grep -oE 'LDGSTS[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|UBLKPF[.A-Z0-9]*|FFMA2|FMUL2|MUFU\.[A-Z0-9]+|LDS\.128|FMNMX\.NAN' profiles/$3 | sort | uniq -c | tr '\n' ' ')"
}
dump fwd_t0_g0.o dhfk_fwd_kernelILb0ELb1ELi0ELb0ELb0E ${TAG}_sass_dhfk_fwd_kernel_cam0_uv1_trig0_gen0.txt
dump bwd_t1_b0_g0.o dhfk_bwd_kernelILb1ELb0ELb1ELb0ELi1ELb0ELb0E ${TAG}_sass_dhfk_bwd_kernel_gw1_gcam0_guv1_gbone0_trig1_gen0.txt
