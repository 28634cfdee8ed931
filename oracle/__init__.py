"""CPU oracle for the SDRainer hot path -- TEST INFRASTRUCTURE ONLY (see sdr_oracle.h)."""
