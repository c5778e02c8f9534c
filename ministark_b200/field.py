"""Host-side field descriptors mirroring reference src/field.rs:9-109 (the arithmetic itself runs on
the device; the host only needs constants, canonical reduction and the domain generator)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class StarkField:
    name: str
    field_id: int
    p: int
    generator: int
    two_adicity: int
    ext_degree: int

    @property
    def modulus_bits(self) -> int:
        return self.p.bit_length()

    @property
    def base_bytes(self) -> int:
        return (self.modulus_bits + 7) // 8

    def root_of_unity(self, log_n: int) -> int:
        """ark-poly Radix2EvaluationDomain::group_gen for size 2^log_n."""
        root = pow(self.generator, (self.p - 1) >> self.two_adicity, self.p)
        return pow(root, 1 << (self.two_adicity - log_n), self.p)


Goldilocks = StarkField("Goldilocks", 0, 2**64 - 2**32 + 1, 7, 32, 2)  # field.rs:43-62
BabyBear = StarkField("BabyBear", 1, 2013265921, 440564289, 27, 4)  # field.rs:72-109
FIELDS = {0: Goldilocks, 1: BabyBear}
