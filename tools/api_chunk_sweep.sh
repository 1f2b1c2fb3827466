B=hydrium_b200/bin/api_bench
echo "default: $($B --reps 20 --warmup 3)"
for b in 8 16 32 64; do for d in 4 8; do echo "batch=$b depth=$d: $(HYDRIUM_B200_BATCH=$b HYDRIUM_B200_DEPTH=$d $B --reps 10 --warmup 3 | cut -c100-230)"; done; done
echo "one-frame: $($B --one-frame --reps 10 --warmup 3 | cut -c100-230)"
echo "shift3: $($B --shift 3 --reps 10 --warmup 3 | cut -c100-230)"
HYDRIUM_B200_APITRACE=1 $B --reps 2 --warmup 2 2>&1 | grep -v "^{" | tail -10
