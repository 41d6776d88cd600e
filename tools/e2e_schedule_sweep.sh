# End-to-end time of lteb200_dl_receive_host on the bench workload for a few chunk schedules (development aid; output kept as
# profiles/r2_e2e_schedule.txt).  LTEB200_E2E_CHUNK = equal chunks, LTEB200_E2E_SCHEDULE = explicit sizes, neither = the shipped wave-aligned ramp.
run() { echo "== $1"; python tools/e2e_trace.py 4000 2>&1 | grep -E "e2e trace|wall ms"; }
run "shipped schedule (wave-aligned ramp 2 -> 12 waves, small last chunk)"
LTEB200_E2E_CHUNK=296 run "equal chunks of 296 subframes (4 waves)"
LTEB200_E2E_CHUNK=592 run "equal chunks of 592 subframes (8 waves)"
LTEB200_E2E_CHUNK=1036 run "equal chunks of 1036 subframes (14 waves)"
LTEB200_E2E_SCHEDULE=128,256,512,1024,1024,768 run "round-1 style doubling 128,256,512,1024,1024,768,288"
LTEB200_E2E_CHUNK=4000 run "one chunk (no pipelining)"
