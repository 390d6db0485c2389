cd /root/repo
timeout 800 python scripts/energy_drift3.py 2>&1 | tail -6 | cut -c1-700
