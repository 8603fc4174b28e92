#!/bin/bash
# --best on the 10-s stereo fixture (seed 2) through the CLI: size, bps, wall time, decode + cmp
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/synth_wav.py /tmp/stereo10.wav 10 2 2
md5sum /tmp/stereo10.wav
for G in ${GENS:-128}; do
  time sac_b200/sac --encode --best --opt-cfg=dds,$G,0.25 /tmp/stereo10.wav /tmp/s10_$G.sac 2>&1 | tail -8
  ls -l /tmp/s10_$G.sac | awk '{print "dds,'$G' bytes", $5, "bps", $5*8/(441000*2)}'
done
[ -n "$NODECODE" ] || time sac_b200/sac --decode /tmp/s10_${GENS##* }.sac /tmp/s10_back.wav 2>&1 | tail -4
[ -n "$NODECODE" ] || cmp /tmp/stereo10.wav /tmp/s10_back.wav && echo "round trip: bit-exact"
