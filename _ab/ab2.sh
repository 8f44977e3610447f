python scripts/run_plan.py c4 --plan c4 --lib _ab/libold.so 2>&1 | tail -1
python scripts/run_plan.py c4 --plan c4 2>&1 | tail -1
python scripts/run_plan.py c4 --plan c4 --tc-pair 5 2>&1 | tail -1
python scripts/run_plan.py c4 --plan c4 --tc-pair 7 2>&1 | tail -1
python scripts/run_plan.py c4 --plan c4 --lib _ab/libold.so 2>&1 | tail -1
python scripts/run_plan.py c4 --plan c4 --tc-pair 5 2>&1 | tail -1
