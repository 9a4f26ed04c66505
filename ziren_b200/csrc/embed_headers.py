#!/usr/bin/env python3
"""file:symbol ... -> C++ raw string literals on stdout (Makefile: embedded_headers.inc)."""
import sys

for arg in sys.argv[1:]:
    path, sym = arg.split(":")
    text = open(path).read()
    assert ')ZKB"' not in text
    # a raw string literal is limited to 64 KB by some front ends: split into adjacent literals
    parts = [text[i:i + 16000] for i in range(0, len(text), 16000)]
    print("static const char %s[] =" % sym)
    for p in parts:
        print('R"ZKB(' + p + ')ZKB"')
    print(";")
