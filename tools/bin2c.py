#!/usr/bin/env python
"""bin2c.py <file> <symbol>  ->  C header on stdout: `static const unsigned char <symbol>[] = {...}; static const size_t <symbol>_len`."""
import sys

data = open(sys.argv[1], "rb").read()
name = sys.argv[2]
print(f"// generated from {sys.argv[1].split('/')[-1]} ({len(data)} bytes) by tools/bin2c.py - do not edit")
print("#pragma once\n#include <cstddef>")
print(f"alignas(64) static const unsigned char {name}[] = {{")
for i in range(0, len(data), 24):
    print("  " + ",".join(str(b) for b in data[i:i + 24]) + ",")
print("};")
print(f"static const size_t {name}_len = {len(data)};")
