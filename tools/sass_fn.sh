#!/bin/sh
# usage: tools/sass_fn.sh <mangled function name> [lib]   -> normalised SASS (addresses stripped) of one kernel on stdout
LIB=${2:-openmines_b200/libopenmines_b200.so}
cuobjdump -sass "$LIB" | awk -v fn="$1" '/Function :/ {on = ($3 == fn)} on && /^ +\/\*[0-9a-f]{4}\*\// {sub(/\/\*[0-9a-f]+\*\//, ""); sub(/\/\* 0x[0-9a-f]+ \*\//, ""); print}'
