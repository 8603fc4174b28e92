#!/bin/bash
# configs[1]: stereo 60-s fixture (seed 3) at --high with the reference's own sequential DDS and warm start:
# size against the reference's measured 6 966 708 bytes, then decode + cmp
set -e
cd "$(dirname "$0")/.."
python tools/synth_wav.py /tmp/stereo60.wav 60 2 3
md5sum /tmp/stereo60.wav
time sac_b200/sac --encode --high --opt-cfg=dds,0 /tmp/stereo60.wav /tmp/s60_high.sac 2>&1 | tail -8
ls -l /tmp/s60_high.sac | awk '{print "--high bytes", $5, "bps", $5*8/(2646000*2), "reference 6966708 -> delta", $5-6966708, "(" ($5-6966708)*100/6966708 " %)"}'
time sac_b200/sac --decode /tmp/s60_high.sac /tmp/s60_back.wav 2>&1 | tail -3
cmp /tmp/stereo60.wav /tmp/s60_back.wav && echo "round trip: bit-exact"
