#!/bin/sh
# usage: tools/sass_fn.sh <mangled function name> [lib]   -> normalised SASS (addresses stripped) of one kernel on stdout
LIB=${2:-openmines_b200/libopenmines_b200.so}
cuobjdump -sass "$LIB" | awk -v fn="$1" '/Function :/ {on = ($3 == fn)} on' | grep -E '^ +/\*[0-9a-f]+\*/ ' | sed -E 's#/\*[0-9a-f]+\*/##; s#/\* 0x[0-9a-f]+ \*/##'
