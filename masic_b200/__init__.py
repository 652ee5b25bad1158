"""masic_b200 — B200-native (sm_100a) hot path of the MASIC stereo image codec.

    from masic_b200.hsic import HSIC          # drop-in for coremasic/mywork/MASIC.py:HSIC
    python -m masic_b200.build                # nvcc -> masic_b200/libmasic_b200.so

See DESIGN.md (kernels, layouts, parity), INTEGRATION.md (C ABI binding), include/masic_b200.h.
"""
__version__ = "0.1.0"
