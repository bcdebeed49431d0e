set -e
cd imhd-cuda_b200/driver
df -h /tmp | tail -1
sed -e 's/^Nt=.*/Nt=201/' -e 's/^Nx=.*/Nx=304/' -e 's/^Ny=.*/Ny=304/' -e 's/^Nz=.*/Nz=592/' input.inp > /tmp/c5.inp
for every in 50 100000; do
  IMHD_OUTPUT_EVERY=$every python simulation_launcher.py nodiff --input /tmp/c5.inp --data-dir /tmp/c5data | tail -1
  ls /tmp/c5data | wc -l; du -sh /tmp/c5data | cut -f1
done
