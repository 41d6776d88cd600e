run() { python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$1', d['e2e']['ms_per_step_all_repeats'], d['e2e']['value'])"; }
run default
export LTEB200_E2E_SCHEDULE=256,512,1024,1184,768; run 256,512,1024,1184,768,256
export LTEB200_E2E_SCHEDULE=192,384,768,1024,1024,416; run 192,384,768,1024,1024,416,192
export LTEB200_E2E_SCHEDULE=160,320,640,1280,1280,224; run 160,320,640,1280,1280,224,96
export LTEB200_E2E_SCHEDULE=128,256,512,1024,1024,768; run 128,256,512,1024,1024,768,288
