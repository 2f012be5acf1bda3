/* The 122,784-byte static dictionary of the Brotli specification (appendix A), embedded from
 * brotli_rs_b200/data/dictionary.bin (generated and CRC-checked by tools/gen_tables.py).
 * Compile with -Wa,-I<dir containing dictionary.bin>.  Reference: src/dictionary/mod.rs:13. */
__asm__(
    ".section .rodata\n"
    ".balign 16\n"
    ".global bro_dictionary_blob\n"
    ".type bro_dictionary_blob, @object\n"
    "bro_dictionary_blob:\n"
    ".incbin \"dictionary.bin\"\n"
    ".size bro_dictionary_blob, .-bro_dictionary_blob\n"
    ".previous\n");
