/* Embeds the float model's weight pack into libvqvdb_b200.so (the reference embeds its
 * model blob the same way, as a C array: src/Bin/bin_model.h:14).  VQVDB_PACK_PATH is
 * supplied by the build (vqvdb_b200/build.py). */
	.section .rodata
	.balign 64
	.global vqvdb_b200_embedded_pack
	.type vqvdb_b200_embedded_pack, @object
vqvdb_b200_embedded_pack:
	.incbin VQVDB_PACK_PATH
	.global vqvdb_b200_embedded_pack_end
	.type vqvdb_b200_embedded_pack_end, @object
vqvdb_b200_embedded_pack_end:
	.byte 0
	.size vqvdb_b200_embedded_pack, vqvdb_b200_embedded_pack_end - vqvdb_b200_embedded_pack
	.section .note.GNU-stack,"",@progbits
