"""Mirror of the reference's `util` module (src/util.rs:4-44): the three integer helpers the tree and the parameter
derivation are written in.  Same names, same results, same failure behaviour (`Err(&str)` -> ValueError with that
string, `assert!` -> AssertionError)."""
from __future__ import annotations


def is_power_of_two(number: int) -> bool:  # util.rs:4-14 (zero counts as a power of two)
    if number > 0:
        return number & (number - 1) == 0
    return number == 0


def _trailing_zeros(v: int, bits: int = 64) -> int:
    return bits if v == 0 else (v & -v).bit_length() - 1


def logarithm_of_two_k(number: int, base: int) -> int:  # util.rs:16-28
    assert is_power_of_two(base)
    log_n = _trailing_zeros(base)
    if not is_power_of_two(number):
        raise ValueError("number if not a power of 2")
    power_of_two = _trailing_zeros(number)
    if power_of_two % log_n != 0:
        raise ValueError("number if not a power of base")
    return power_of_two // log_n


def ceil_log2_k(number: int, base: int) -> int:  # util.rs:30-44
    assert is_power_of_two(base)
    assert number != 0
    if number == 1:
        return 1
    log2_base = _trailing_zeros(base)
    log2_number = _trailing_zeros(number)
    if is_power_of_two(number) and log2_number % log2_base == 0:
        return log2_number
    next_power_2 = number.bit_length()  # usize::BITS - leading_zeros
    return -(-next_power_2 // log2_base) * log2_base
