#!/bin/bash
# Regenerates the CLI fixture: a 0.25-s stereo WAV (tools/synth_wav.py seed 9) with an odd-sized LIST chunk, encoded by the
# UNMODIFIED reference CLI (oracle/_ref/sac, built by oracle/Makefile) at --normal, plus the reference's own --list / --listfull
# output for it. tests/test_host_logic.py compares our CLI's listing with these, line for line. Needs /root/reference (via oracle/_ref).
set -e
here="$(cd "$(dirname "$0")" && pwd)"; root="$here/../.."
tmp=$(mktemp -d)
python - "$tmp/q.wav" <<'PY'
import sys, struct
sys.path.insert(0, "tools")
from synth_wav import synth_pcm
pcm = synth_pcm(0.25, 2, 9); data = pcm.astype('<i2').tobytes()
fmt = struct.pack('<HHIIHH', 1, 2, 44100, 44100 * 4, 4, 16)
body = b'WAVE' + b'fmt ' + struct.pack('<I', 16) + fmt + b'LIST' + struct.pack('<I', 7) + b'INFOabc\0' + b'data' + struct.pack('<I', len(data)) + data
open(sys.argv[1], 'wb').write(b'RIFF' + struct.pack('<I', len(body)) + body)
PY
cd "$tmp"
"$root/oracle/_ref/sac" --encode --normal q.wav q.sac > /dev/null
cp q.sac "$here/ref_stereo_quarter_normal.sac"
cd "$here"
"$root/oracle/_ref/sac" --listfull ref_stereo_quarter_normal.sac | sed -n '/^Open/,$p' > ref_stereo_quarter_normal.listfull.txt
"$root/oracle/_ref/sac" --list ref_stereo_quarter_normal.sac | sed -n '/^Open/,$p' > ref_stereo_quarter_normal.list.txt

# Second fixture: a 0.2-s mono WAV whose samples sit on a 16-step grid (12-bit audio in a 16-bit container,
# make_golden_sparse.py "shift4" seed 61): the reference codes its frame rank-mapped (block flag 1<<9 | maxbpn_map).
python - "$tmp/s.wav" <<'PY'
import sys, struct
sys.path.insert(0, "tools"); sys.path.insert(0, "tests/golden"); sys.path.insert(0, "tests")
from make_golden_sparse import sparse_pcm
pcm = sparse_pcm("shift4", 0.2, 1, 61); data = pcm.astype('<i2').tobytes()
fmt = struct.pack('<HHIIHH', 1, 1, 44100, 44100 * 2, 2, 16)
body = b'WAVE' + b'fmt ' + struct.pack('<I', 16) + fmt + b'data' + struct.pack('<I', len(data)) + data
open(sys.argv[1], 'wb').write(b'RIFF' + struct.pack('<I', len(body)) + body)
PY
cd "$tmp"
"$root/oracle/_ref/sac" --encode --normal s.wav s.sac > /dev/null
cp s.sac "$here/ref_mono_sparse_normal.sac"
cd "$here"
"$root/oracle/_ref/sac" --listfull ref_mono_sparse_normal.sac | sed -n '/^Open/,$p' > ref_mono_sparse_normal.listfull.txt
