/* Embeds weights/acnet.bin (the reference's ACNet tables, produced by tools/gen_weights.cpp) into the library. */
    .section .rodata
    .global acb200_weights_blob
    .global acb200_weights_blob_end
    .balign 64
acb200_weights_blob:
    .incbin AC_WEIGHTS_BIN
acb200_weights_blob_end:
    .byte 0
    .section .note.GNU-stack,"",@progbits
