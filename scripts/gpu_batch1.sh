# one GPU slot: profile-build counters of the inverse kernel, then (normal build is NOT available in this call)
timeout 120 python scripts/prof_made_inverse.py 4736
