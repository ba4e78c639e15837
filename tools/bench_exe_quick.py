#!/usr/bin/env python
"""Development aid: the text-in / text-out leg of bench.py alone (paladin-gpu --right on MatrixMarket files of 512x512)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


class A:
    exe_matrices = int(sys.argv[1]) if len(sys.argv) > 1 else 1184


print(json.dumps(bench.exe_entry(None, A())))
