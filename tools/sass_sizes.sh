#!/bin/bash
# static SASS instruction count per kernel of a built library: tools/sass_sizes.sh [lib.so]
cuobjdump -sass "${1:-akari_render_b200/libakari_b200.so}" 2>/dev/null | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/{c[name]++} END{for(n in c) print c[n], n}' | sort -rn | c++filt | sed 's/(anonymous namespace):://g; s/(.*//' | head -${2:-24}
