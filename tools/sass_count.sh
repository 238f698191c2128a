#!/bin/bash
# SASS instruction counts of the headline kernels of a built library: tools/sass_count.sh [LIB.so]
LIB=${1:-openmines_b200/libopenmines_b200.so}
for k in _Z13om_run_kernelILi1ELb0ELb0ELb0EEv9RunParams _Z13om_run_kernelILi0ELb0ELb0ELb0EEv9RunParams _Z13om_run_kernelILi1ELb0ELb1ELb0EEv9RunParams _Z14om_step_kernelILi1ELb0EEv9RunParams; do
  n=$(cuobjdump -sass "$LIB" | awk -v k="$k" '/Function : /{on=($3==k)} on && /^ +\/\*[0-9a-f]+\*\/ /{n++} END{print n}')
  echo "$k $n"
done
