#!/bin/bash
#SBATCH --job-name=tdp_b200
#SBATCH --nodes=2
#SBATCH --ntasks-per-node=8
#SBATCH --gres=gpu:8
#SBATCH --cpus-per-task=12
# one task per GPU; setup_distributed() reads the SLURM_* variables
srun python -m torchdistpackage_b200.dist.py_comm_test --mode all_reduce --mib 1024
