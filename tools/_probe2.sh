python tools/step_profile.py 2>&1 | grep -E "gmm|total" > gpurun_out/sp_base.txt
MASIC_GMM_L1_CG2=1 MASIC_GMM_L2_CG2=1 python tools/step_profile.py 2>&1 | grep -E "gmm|total" > gpurun_out/sp_cg2.txt
paste gpurun_out/sp_base.txt gpurun_out/sp_cg2.txt | cut -c1-250
