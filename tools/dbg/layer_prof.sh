set -u
O=gpurun_out
for T in l32_64 l32_1 l16_32 l16_16; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_t -s 2 -c 1 -f -o $O/prof_$T python tools/prof_kernel.py $T 3 > $O/ncu_$T.log 2>&1
done
python tools/make_profiles.py r02 > $O/make_profiles.log 2>&1
rm -rf $O/profiles_new && cp -r profiles $O/profiles_new
rm -f $O/*.ncu-rep
