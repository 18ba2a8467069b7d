#!/bin/bash
# CUPTI timeline of one device-resident C2 step with and without the one-launch set-up
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for x in 0 1; do echo "== URE_SETUP_FUSED=$x"; URE_SETUP_FUSED=$x timeout 600 python tools/prof_timeline.py 50 2>&1 | head -45; done
