"""Thin one-frame images (LF groups a few pixels high or wide) against the reference library: the shapes
that exercised the parallel LFGroup coder's corner cases."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrium_b200.encoder import encode_cli_loop
from hydrium_b200.lib import load_library
from hydrium_b200.synth import synth_image
from oracle.pyoracle import ref_library
lib=load_library(); ref=ref_library()
for (w,h,smooth) in [(2048,9,True),(2048,8,True),(2048,9,False),(512,9,True),(600,16,True),(2048,64,True),(2048,300,True),(1024,24,True)]:
    img=synth_image(w,h,8,seed=6,smooth=smooth)
    a=encode_cli_loop(lib,img,shift_x=-1,shift_y=-1); b=encode_cli_loop(ref,img,shift_x=-1,shift_y=-1)
    n=min(len(a),len(b)); first=next((i for i in range(n) if a[i]!=b[i]), n)
    print(w,h,smooth,len(a),len(b),"OK" if a==b else f"DIFF at {first}")
