/* Embeds the sm_100a cubin into libaule.so (analogue of @embedFile in src/lib.zig:29-50). */
    .section .rodata
    .global aule_cubin_start
    .global aule_cubin_end
    .balign 64
aule_cubin_start:
    .incbin AULE_CUBIN_PATH
aule_cubin_end:
    .byte 0
    .section .note.GNU-stack,"",@progbits
