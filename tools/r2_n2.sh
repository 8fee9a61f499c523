# round 2: two ranks -- bench through md_comm_init/md_gather_psms (NCCL inside the library), clean exit; C++ host with --nranks 2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -8 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('value',d['value'],'e2e',d['e2e'],'ms/step',d['ms_per_step'],'crc',d['psm_crc'],'rows',d['psm_rows'])
print('c4',d.get('c4_strong'))
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
# C++ host: one rank vs two ranks
python - <<'PY'
import sys; sys.path.insert(0,'max-decoy_b200')
from maxdecoy import synth
pr=synth.synthetic_proteins(300)
open('/tmp/db.fasta','w').write(synth.fasta_text(pr))
sp,_=synth.synthetic_spectra(pr,200,2,mods=(synth.CAM,synth.OXM))
open('/tmp/run.mgf','w').write(synth.mgf_text(sp))
open('/tmp/mods.csv','w').write(synth.mods_csv_text([synth.CAM,synth.OXM]))
PY
H=max-decoy_b200/host/max_decoy
COMMON="identification -m /tmp/mods.csv -s /tmp/run.mgf --fasta /tmp/db.fasta -n 3 -d 50 -l 10 -u 10 --seed 5"
$H $COMMON -o /tmp/out1 ; echo "host 1 rank rc=$?"
rm -f /tmp/comm.id
$H $COMMON -o /tmp/out2 --rank 1 --nranks 2 --comm-file /tmp/comm.id --device 1 &
$H $COMMON -o /tmp/out2 --rank 0 --nranks 2 --comm-file /tmp/comm.id --device 0 ; echo "host rank0 rc=$?"
wait; echo "host rank1 rc=$?"
cmp /tmp/out1/psms.csv /tmp/out2/psms.csv && echo "psms.csv identical (1 rank vs 2 ranks)"; wc -l /tmp/out1/psms.csv /tmp/out2/psms.csv
diff <(cd /tmp/out1; ls | wc -l) <(cd /tmp/out2; ls | wc -l) && echo "same number of files"
for f in $(cd /tmp/out1; ls *.fasta | head -40); do cmp -s /tmp/out1/$f /tmp/out2/$f || echo "DIFF $f"; done; echo "fasta compare done"
