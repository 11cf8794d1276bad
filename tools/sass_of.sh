#!/bin/bash
# usage: tools/sass_of.sh <substring of the mangled kernel name> [lib]   -> "addr instruction" lines of that kernel
LIB=${2:-jetson_slam_b200/libjsfe.so}
cuobjdump -sass "$LIB" | awk -v k="$1" '/Function : /{f=index($0,k)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
