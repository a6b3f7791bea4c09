#pragma once
#include <cstddef>
#include <cstdint>
namespace tfh {
// rfc = false (default): tag as the reference computes it (S:192-270, a non-standard limb recombination);
// rfc = true: RFC 8439 tag.
void aead_seal(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen, uint8_t* data, size_t n, uint8_t tag[16], bool rfc = false);
bool aead_open(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen, uint8_t* data, size_t n, const uint8_t tag[16], bool rfc = false);
}  // namespace tfh
