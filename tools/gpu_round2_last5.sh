# The last ~70 GPU-seconds: host cost of a part-step without the Python face and without the caller-stream fence.
mkdir -p gpurun_out
( timeout -s KILL 25 python tools/halves_timeline.py deep_sea/11 65536 raw nofence 4
  timeout -s KILL 25 python tools/halves_timeline.py deep_sea/11 65536 raw fence 4
  timeout -s KILL 25 python tools/halves_timeline.py deep_sea/11 65536 raw nofence 3 ) > gpurun_out/last5_raw_timeline.txt 2>&1
cat gpurun_out/last5_raw_timeline.txt
